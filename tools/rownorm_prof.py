"""Wait-cycle profile of dn_gemm_resid_norm (debugging aid).  Needs the instrumented build of that one file:
    cd diffnorm_b200/csrc && nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
        -DRN_PROFILE -c gemm_rownorm.cu -o gemm_rownorm.o && make      # relinks; `touch gemm_rownorm.cu && make` undoes it
Output of the round-1 runs: profiles/r01_rn_resid_norm_wait_cycles.txt."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffnorm_b200 import _lib  # noqa: E402
from tools.rownorm_check import case, fused  # noqa: E402

NAMES = ["producer wait empty", "mma wait tempty", "mma wait full", "x-thread wait xready", "x-thread wait store-read",
         "hb-thread wait hready", "hb-thread wait store-read", "pass1 wait tfull", "pass1 wait xfull", "pass2 wait ssfull",
         "pass2 wait hfree", "kernel cycles", "pass1 tmem_ld wait", "pass1 compute", "pass1 fence.proxy", "pass1 tmem_st+wait"]
fn = _lib.lib.dn_debug_rownorm_profile
fn.argtypes = [C.c_void_p]
out = (C.c_ulonglong * 16)()
dev = torch.device("cuda:0")
for K in (512, 1365):
    plan, A, x0, kw = case(64000, K, K != 512, True)
    x = x0.clone()
    hb = torch.zeros(64000, 512, dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        fused(plan, A, x, hb, kw)
    fn(out)
    iters = 5
    for _ in range(iters):
        fused(plan, A, x, hb, kw)
    fn(out)
    ctas = 148
    tot = out[11] / ctas / iters
    print(f"K={K}: kernel {tot:.0f} cycles per CTA")
    for i, n in enumerate(NAMES):
        v = out[i] / ctas / iters
        print(f"   {n:28s} {v:9.0f}  {100 * v / tot:5.1f} %")
