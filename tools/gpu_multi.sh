#!/bin/bash
# usage: tools/gpu_multi.sh TAG N [stages...]   stages: pass train sweep dataset
TAG=$1; N=$2; shift 2; STAGES=${*:-pass train dataset}
O=gpurun_out; mkdir -p $O
has() { [[ " $STAGES " == *" $1 "* ]]; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "1" ]; then TR="python"; fi
if has pass; then timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 > $O/${TAG}_pass$N.log 2>&1; echo "pass $?"; fi
if has train; then timeout 600 $TR bench.py --gpus $N --config train --steps 5 --warmup 3 > $O/${TAG}_train$N.log 2>&1; echo "train $?"; fi
if has sweep; then timeout 900 $TR bench.py --gpus $N --config train --steps 5 --warmup 3 --comm-sweep > $O/${TAG}_trainsweep$N.log 2>&1; echo "sweep $?"; fi
if has dataset; then timeout 1500 $TR bench.py --gpus $N --config dataset --utts ${UTTS:-20000} > $O/${TAG}_dataset$N.log 2>&1; echo "dataset $?"; fi
