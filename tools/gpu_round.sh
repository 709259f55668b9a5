#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, HBM-kernel bandwidths, config-3/4 sweeps, then (only after the
# plain runs exited 0) the ncu launch list and one --set full capture of the two top kernels.  Output -> gpurun_out/.
# usage: tools/gpu_round.sh TAG [stages...]   stages: tests bench hbm sweep dataset break ncu
set -u
TAG=${1:-t}
shift || true
STAGES=${*:-tests bench hbm sweep dataset break ncu}
O=gpurun_out
mkdir -p $O
: > $O/${TAG}_info.log
has() { [[ " $STAGES " == *" $1 "* ]]; }
if has tests; then
  timeout 900 python -m pytest tests -m gpu -x -q -s > $O/${TAG}_tests.log 2>&1; echo "tests exit $?" >> $O/${TAG}_info.log
fi
if has bench; then
  timeout 600 python bench.py --steps 3 --warmup 3 > $O/${TAG}_bench.log 2>&1; echo "bench exit $?" >> $O/${TAG}_info.log
fi
if has break; then
  timeout 300 python tools/breakdown.py > $O/${TAG}_break.log 2>&1; echo "break exit $?" >> $O/${TAG}_info.log
  timeout 300 python tools/breakdown.py --what decode > $O/${TAG}_break_decode.log 2>&1
fi
if has hbm; then
  timeout 300 python tools/hbm_kernels.py > $O/${TAG}_hbm.log 2>&1; echo "hbm exit $?" >> $O/${TAG}_info.log
fi
if has sweep; then
  timeout 900 python tools/sweep.py --what sweep > $O/${TAG}_sweep.log 2>&1; echo "sweep exit $?" >> $O/${TAG}_info.log
fi
if has dataset; then
  timeout 900 python tools/sweep.py --what dataset --utts 2000 > $O/${TAG}_dataset.log 2>&1; echo "dataset exit $?" >> $O/${TAG}_info.log
fi
if has ncu; then
  # launch list of a short pass (start_step 3 = 2 denoiser calls; same kernels, fewer launches)
  timeout 600 python bench.py --steps 1 --warmup 3 --start-step 3 --no-cpu-baseline > $O/${TAG}_plain3.log 2>&1
  rc=$?; echo "plain3 exit $rc" >> $O/${TAG}_info.log
  if [ $rc -eq 0 ]; then
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches.csv \
      python bench.py --steps 1 --warmup 3 --start-step 3 --no-cpu-baseline > $O/${TAG}_ncu_launches.log 2>&1
    echo "ncu launches exit $?" >> $O/${TAG}_info.log
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 2 -c 1 \
      -o $O/${TAG}_attn_tc -f python tools/breakdown.py > $O/${TAG}_ncu_attn.log 2>&1
    echo "ncu attn exit $?" >> $O/${TAG}_info.log
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 8 \
      -o $O/${TAG}_gemm_layer -f python tools/breakdown.py > $O/${TAG}_ncu_gemm.log 2>&1
    echo "ncu gemm exit $?" >> $O/${TAG}_info.log
  fi
fi
cat $O/${TAG}_info.log
