"""World-size-2 gloo test of the gradient exchange used by the denoiser training step (the only collective on any
path): bucketing in production order, async launch, mean reduction, views handed back per parameter."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffnorm_b200.dist import GradAllReducer
    ok = True
    # fp32 wire format (DDP's arithmetic), launched as buckets fill or all at the end; bf16 wire format (values here are
    # small integers and halves: exact in bf16)
    for kw in (dict(), dict(overlap=False), dict(comm_dtype=torch.bfloat16)):
        red = GradAllReducer(bucket_bytes=4 * 100, **kw)   # ~100 floats per bucket -> several buckets, one oversize tensor
        for step in range(2):                        # buckets are persistent across steps
            shapes = [(7, 5), (64,), (3, 4, 3), (250,), (1,), (30, 3)]
            grads = {f"p{i}": torch.full(s, float((rank + 1) * (i + 1) + step)) for i, s in enumerate(shapes)}
            for k, v in grads.items():
                red.hook(k, v)
            out = red.finish()
            for i, s in enumerate(shapes):
                want = sum((r + 1) * (i + 1) + step for r in range(world)) / world
                ok &= out[f"p{i}"].shape == torch.Size(s) and out[f"p{i}"].dtype == torch.float32
                ok &= bool(torch.allclose(out[f"p{i}"], torch.full(s, want)))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_grad_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_grad_allreduce_single_process_is_identity():
    from diffnorm_b200.dist import GradAllReducer
    red = GradAllReducer(bucket_bytes=64)
    g = {"a": torch.arange(10.0), "b": torch.ones(3, 3)}
    for k, v in g.items():
        red.hook(k, v)
    out = red.finish()
    assert torch.equal(out["a"], g["a"]) and torch.equal(out["b"], g["b"])
