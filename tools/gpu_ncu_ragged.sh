#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gemm_tc_kernel<\(int\)0, \(int\)2, \(int\)0>' -s 10 -c 4 -o $O/z2_conv_ragged \
  python tools/breakdown.py --batch 106 --frames 600 > $O/z2_ncu_conv.log 2>&1; echo "ncu full $?"
ls -la $O/z2_conv_ragged.ncu-rep
