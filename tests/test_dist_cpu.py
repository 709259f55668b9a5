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


# ---- the normalization path at N > 1: sharded by utterance, no data-path collective; the one cross-rank step is the host-side
# ---- stitching of the per-rank TSV shards (normalize_cli.write_shards)
def _norm_worker(rank, world, port, out_dir, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    from diffnorm_b200.data import plan_batches
    from diffnorm_b200.normalize_cli import write_shards
    rng = np.random.default_rng(7)          # every rank derives the same plan from the same lengths: nothing is exchanged
    lengths = np.clip(np.rint(np.exp(rng.normal(np.log(600.0), 0.5, size=301))), 200, 2000).astype(np.int64)
    plan = plan_batches(lengths, 16000, world_size=world)
    mine = np.concatenate(plan[rank])
    lines = {int(i): f"utt{int(i)}\t{int(lengths[i])}" for i in mine}      # stands for this rank's normalized TSV lines
    path = write_shards(lines, out_dir, "train", rank, world)
    loads = [int(sum(len(b) * int(lengths[b].max()) for b in plan[r])) for r in range(world)]
    q.put((rank, len(mine), os.path.basename(path), loads))
    dist.barrier()
    dist.destroy_process_group()


def test_normalization_sharding_and_stitching_world2_gloo(tmp_path):
    import numpy as np
    world, port = 2, 29533
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_norm_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p_ in ps:
        p_.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p_ in ps:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    assert sum(r[1] for r in res) == 301 and res[0][2] == "train.tsv" and res[1][2] == "train.rank1.tsv"
    loads = res[0][3]
    assert max(loads) <= 1.15 * min(loads)                      # LPT plan on ~13 batches per rank (its cost model also weighs T^2)
    rng = np.random.default_rng(7)
    lengths = np.clip(np.rint(np.exp(rng.normal(np.log(600.0), 0.5, size=301))), 200, 2000).astype(np.int64)
    merged = open(os.path.join(str(tmp_path), "train.tsv")).read().splitlines()[1:]
    assert merged == [f"utt{i}\t{int(lengths[i])}" for i in range(301)]     # = the file one process writes, original order
    shards = [open(os.path.join(str(tmp_path), f"train.rank{r}.tsv")).read().splitlines()[1:] for r in range(world)]
    assert sorted(shards[0] + shards[1]) == sorted(merged) and not set(shards[0]) & set(shards[1])
