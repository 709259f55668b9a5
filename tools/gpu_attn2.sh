O=gpurun_out
for s in 0 400 900 1400 2000; do F16=0 DN_ATTN_STAGGER_NS=$s timeout 60 python tools/attn_bench.py >> $O/p2_attn.log 2>&1; echo "stagger $s" >> $O/p2_attn.log; done
F16=0 DN_ATTN_PERSIST=0 timeout 60 python tools/attn_bench.py >> $O/p2_attn.log 2>&1
