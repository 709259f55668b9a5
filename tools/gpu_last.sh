#!/bin/bash
# last check of the committed build: kernel + pass + runner tests, smoke, default bench
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_pass_gpu.py tests/test_runner_gpu.py -q --timeout=300 > $O/z3_tests.log 2>&1; echo "tests $?"; tail -2 $O/z3_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/z3_smoke.log 2>&1; echo "smoke $?"; tail -1 $O/z3_smoke.log
timeout 600 python bench.py --no-cpu-baseline > $O/z3_bench.log 2>&1; echo "bench $?"; tail -1 $O/z3_bench.log | cut -c1-160
