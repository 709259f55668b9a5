O=gpurun_out
timeout 200 python -m pytest tests/test_precision_gpu.py -q -x --timeout=100 -k "stochastic" > $O/s1_test.log 2>&1; echo "t $?"
DN_WFMT=bf16sr timeout 300 python tools/parity_fullsize.py > $O/s1_parity_sr.json 2> $O/s1_parity_sr.err; echo "p $?"
DN_WFMT=bf16sr timeout 300 python tools/parity_fullsize.py --batch 8 --frames 1000 --chunk 4 > $O/s1_parity_sr_t1000.json 2>> $O/s1_parity_sr.err; echo "p2 $?"
for i in 1 2; do
  DN_WFMT=bf16sr timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/s1_bench_sr_$i.log 2>&1
  DN_WFMT=f16 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/s1_bench_f16_$i.log 2>&1
done
DN_WFMT=bf16 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/s1_bench_bf16.log 2>&1
