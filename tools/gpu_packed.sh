#!/bin/bash
# packed-rows GEMM tiling: parity tests, then the denoiser-call breakdown at a ragged and at the uniform shape per chunk size
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x --timeout=120 -k "gemm" > $O/q1_tests.log 2>&1; echo "tests $?"
tail -3 $O/q1_tests.log
for rc in 0 32 16 8; do
  for shape in "106 600" "64 1000" "160 400"; do
    set -- $shape
    DN_ROW_CHUNK=$rc timeout 200 python tools/breakdown.py --batch $1 --frames $2 > $O/q1_bd_rc${rc}_B$1_T$2.txt 2>&1
    echo "rc=$rc B=$1 T=$2: $(head -1 $O/q1_bd_rc${rc}_B$1_T$2.txt)"
  done
done
