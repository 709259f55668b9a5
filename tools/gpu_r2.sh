#!/bin/bash
# round-2 GPU batch: usage tools/gpu_r2.sh TAG [stages...]; stages: prec parity full tests bench break
set -u
TAG=${1:-b}; shift || true
STAGES=${*:-prec parity full}
O=gpurun_out; mkdir -p $O; : > $O/${TAG}_info.log
has() { [[ " $STAGES " == *" $1 "* ]]; }
if has prec; then timeout 600 python -m pytest tests/test_precision_gpu.py -q -s --timeout=150 > $O/${TAG}_prec.log 2>&1; echo "prec exit $?" >> $O/${TAG}_info.log; fi
if has parity; then timeout 600 python tools/parity_fullsize.py > $O/${TAG}_parity.json 2> $O/${TAG}_parity.err; echo "parity exit $?" >> $O/${TAG}_info.log; fi
if has full; then timeout 900 python -m pytest tests/test_fullsize_parity_gpu.py -q -s --timeout=300 > $O/${TAG}_full.log 2>&1; echo "full exit $?" >> $O/${TAG}_info.log; fi
if has tests; then timeout 1500 python -m pytest tests -m gpu -q -s --timeout=300 > $O/${TAG}_tests.log 2>&1; echo "tests exit $?" >> $O/${TAG}_info.log; fi
if has bench; then timeout 600 python bench.py --steps 3 --warmup 3 > $O/${TAG}_bench.log 2>&1; echo "bench exit $?" >> $O/${TAG}_info.log; fi
if has break; then timeout 300 python tools/breakdown.py > $O/${TAG}_break.log 2>&1; echo "break exit $?" >> $O/${TAG}_info.log
  timeout 300 python tools/breakdown.py --what decode > $O/${TAG}_break_decode.log 2>&1; fi
cat $O/${TAG}_info.log
