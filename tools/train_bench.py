#!/usr/bin/env python
"""BASELINE.json configs[4]: denoiser training step (latent-noise MSE, forward + backward) with NCCL gradient
all-reduce over NVLink at 1/2/4/8 B200 (torchrun for > 1 GPU; one process per GPU, data parallel over utterances).

A step = DenoiserTrainer.step (frozen-VAE encode, per-step weight packing, denoiser forward with dropout, the decode
branch's logging losses, full backward) + bucketed async all-reduce of the 260 M fp32 gradients (mean).  The optimizer
belongs to fairseq (out of scope) and is not part of config 5.  Default shape = scripts/diffusion/train.sh's
--max-tokens 12000 per GPU: 12 utterances x 1000 frames.  Prints one JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from diffnorm_b200 import _lib  # noqa: E402
from diffnorm_b200.dist import GradAllReducer  # noqa: E402
from diffnorm_b200.plugin.latent_module import LatentDiscreteModel, SpeechVAEEncoderDecoder  # noqa: E402
from diffnorm_b200.train import DenoiserTrainer  # noqa: E402


def train_flops_per_frame(z, n):
    venc = {16: 1_660_928, 32: 1_626_112, 128: 2_424_832}[z]
    vw = {16: 18_144_256, 32: 18_008_064, 128: 16_809_984}[z]
    d = 141_780_996 + 1024 * z + 12_288 * n
    vdec = vw + 118_554_624 + 771_072 + 9_216 * n
    # forward + data-gradient + weight-gradient of the denoiser (attention backward = 2.5x its forward), VAE forward only
    return 2.0 * (3 * (d - 12_288 * n) + 3.5 * 12_288 * n + venc + vdec)


def vae_bench(a, vae, dev, rank, world, red):
    """VAE training step: forward, the criterion's arithmetic on the returned logits (plain torch, as fairseq's criterion
    does), CUDA backward of all 274 tensors, gradient all-reduce."""
    import torch.nn.functional as F
    from diffnorm_b200.train_vae import VaeTrainer
    B, T = a.batch, a.frames
    tr = VaeTrainer(vae, drop_p=0.1, seed=rank)
    g = torch.Generator().manual_seed(1234 + rank)
    audio = torch.randn(B, T, 768, generator=g).to(dev)
    units = (torch.randint(0, 1000, (B, T), generator=g) + 4).to(dev).view(-1)
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    ntok = B * T

    def step():
        mse, logits, kl = tr.forward(audio, lens)
        lg = logits.detach().requires_grad_(True)
        lp = F.log_softmax(lg, dim=-1).view(-1, lg.shape[-1])
        nll = -lp.gather(1, units[:, None]).sum()
        smooth = -lp.sum()
        e = 0.1 / (lp.shape[-1] - 1)
        (0.1 * ((1 - 0.1 - e) * nll + e * smooth) / ntok).backward()
        tr.backward(10.0, lg.grad, 1e-4, grad_hook=red.hook)
        return mse, red.finish()

    for _ in range(a.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        mse, grads = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        print(json.dumps({"config": "vae_train_step", "metric": "training frames/sec", "value": B * T * world / (ms * 1e-3),
                          "unit": "frames/s", "n_gpus": world, "steps": a.steps, "ms_per_step": ms, "mse": float(mse),
                          "workload": f"VAE training step, {B} x {T} frames per GPU, z {a.latent_dim}, dropout 0.1, fwd + bwd + grad all-reduce",
                          "grad_bytes_allreduced": sum(v.numel() for v in grads.values()) * 4}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=12)
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--latent-dim", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--phases", action="store_true", help="print host/device time of packing, forward, backward (1 GPU)")
    ap.add_argument("--model", default="denoiser", choices=["denoiser", "vae"],
                    help="vae: the VAE training step (scripts/vae/train.sh, --max-tokens 15000: use --batch 15)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    z, B, T = a.latent_dim, a.batch, a.frames
    vae = types.SimpleNamespace(encoder=SpeechVAEEncoderDecoder(768, z))
    ldm = LatentDiscreteModel(vae, 512, z, timesteps=200, multitask=False).to(dev).train()
    tr = DenoiserTrainer(ldm, drop_p=0.1, seed=rank)
    red = GradAllReducer()
    if a.model == "vae":
        return vae_bench(a, ldm.speech_decoder, dev, rank, world, red)
    g = torch.Generator().manual_seed(1234 + rank)
    audio = torch.randn(B, T, 768, generator=g).to(dev)
    units = (torch.randint(0, 1000, (B, T), generator=g) + 4).to(dev)
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)

    def step():
        out, _ = tr.step(audio, units, lens, grad_hook=red.hook)
        grads = red.finish()
        return out, grads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        out, grads = step()
    barrier()
    if a.phases and rank == 0:
        import time

        def timed(fn, n=3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            return (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3

        print("phase: host-enqueue ms / wall ms (3 runs each)")
        print("  pack weights      %.1f / %.1f   (packing.* torch indexing, eager)" % timed(tr._pack))
        print("  per-step refresh  %.2f / %.2f   (DN_REPACK=%s)" % (*timed(tr._packed), os.environ.get("DN_REPACK", "kernel")))
        print("  forward only      %.1f / %.1f" % timed(lambda: tr.step(audio, units, lens, backward=False)))
        print("  forward+backward  %.1f / %.1f" % timed(lambda: tr.step(audio, units, lens)))
        print("  + grad allreduce  %.1f / %.1f" % timed(step))
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out, grads = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / a.steps
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        fpf = train_flops_per_frame(z, T)
        frames = B * T * world
        nbytes = sum(v.numel() for v in grads.values()) * 4
        print(json.dumps({
            "config": "train_step", "metric": "training frames/sec", "value": frames / (ms * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "scaling": "weak", "dtype": "bf16",
            "workload": f"denoiser training step, {B} x {T} frames per GPU, z {z}, dropout 0.1, fwd + bwd + grad all-reduce",
            "loss": float(out["total_loss"]), "grad_bytes_allreduced": nbytes,
            "model_tflops_per_gpu": fpf * B * T / (ms * 1e-3) / 1e12,
            "gpu_launches_per_step": (_lib.launch_count() - n0) // a.steps}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
