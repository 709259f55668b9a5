#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_train_gpu.py -q -x --timeout=200 > $O/q2_train_tests.log 2>&1; echo "train tests $?"; tail -3 $O/q2_train_tests.log
timeout 300 python bench.py --config train --steps 5 --warmup 3 > $O/q2_train1.log 2>&1; echo "train $?"; tail -1 $O/q2_train1.log | cut -c1-300
timeout 900 python bench.py --config dataset --utts 20000 > $O/q2_dataset1.log 2>&1; echo "dataset $?"; tail -1 $O/q2_dataset1.log | cut -c1-300
