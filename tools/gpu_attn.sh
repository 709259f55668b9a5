O=gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_precision_gpu.py -q -x --timeout=60 -k "attention or randn" > $O/p1_tests.log 2>&1; echo "t $?"
for o in 1 0 1 0; do F16=0 DN_ATTN_PERSIST=$o timeout 60 python tools/attn_bench.py >> $O/p1_attn.log 2>&1; done
for o in 1 0; do DN_ATTN_PERSIST=$o timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/p1_bench_persist$o.log 2>&1; done
timeout 300 python -m pytest tests/test_fullsize_parity_gpu.py tests/test_runner_gpu.py -q -x --timeout=200 -k "c1 or long or runner" > $O/p1_full.log 2>&1; echo "f $?"
