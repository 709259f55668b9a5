O=gpurun_out
A8=$PWD/diffnorm_b200/csrc/libdiffnorm_b200_act8.so
timeout 300 python bench.py --config dataset --utts 400 --steps 1 --warmup 1 > $O/c2_dataset400.log 2>&1; echo "dataset $?"
timeout 300 python bench.py --config train --steps 3 --warmup 2 > $O/c2_train1.log 2>&1; echo "train $?"
for i in 1 2; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/c2_bench_f16_$i.log 2>&1
  DN_LIB=$A8 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/c2_bench_act8_$i.log 2>&1
done
DN_LIB=$A8 timeout 300 python tools/parity_fullsize.py > $O/c2_parity_act8.json 2> $O/c2_parity_act8.err; echo "parity act8 $?"
for p in bf16 tf32 fp32; do timeout 400 python bench.py --impl reference-gpu --ref-precision $p --steps 1 > $O/c2_refgpu_$p.log 2>&1; echo "refgpu $p $?"; done
