#!/bin/bash
# round-end verification on the final build: full GPU suite, bench, smoke, ncu launch list of the bench command, one ncu --set full
# capture of the FFN conv GEMM on a ragged batch (packed rows)
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout=300 > $O/z1_tests.log 2>&1; echo "tests $?"; tail -2 $O/z1_tests.log
timeout 600 python bench.py > $O/z1_bench.log 2>&1; echo "bench $?"; tail -1 $O/z1_bench.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/z1_smoke.log 2>&1; echo "smoke $?"; tail -2 $O/z1_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/z1_launches.csv \
  python bench.py --steps 1 --warmup 3 --start-step 3 --no-cpu-baseline > $O/z1_ncu_bench.log 2>&1; echo "ncu list $?"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gemm_tc_kernel<0, 2, 0>' -s 10 -c 4 -o $O/z1_conv_ragged \
  python tools/breakdown.py --batch 106 --frames 600 > $O/z1_ncu_conv.log 2>&1; echo "ncu full $?"
gzip -f $O/z1_launches.csv
