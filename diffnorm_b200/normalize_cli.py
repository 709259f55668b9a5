"""Command-line driver of the normalization pass: drop-in for research/TranSpeech/diff_norm_synthesis.py
(same flags, same input/output file formats, `cli_main` :228-243), with length-bucketed batches and optional
sharding over the GPUs of one node (one process per GPU, e.g. under torchrun; no collective on this path).

    python -m diffnorm_b200.normalize_cli --reduce_tsv_dir R --orig_tsv_dir O --feature_dir F \
        --model_ckpt diffusion.pt --start_step 100 --output_dir OUT [--max-tokens 64000]
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

from .data import NormalizationRunner, prepare_data, write_tsv


def load_model(ckpt_path: str, device: str):
    """checkpoint_utils.load_model_ensemble_and_task for this plugin: torch.load -> setup_task -> build_model(
    from_checkpoint=True) -> load_state_dict(strict=True) (fairseq/checkpoint_utils.py:391-493)."""
    from .plugin import compat
    state = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    args = state.get("args")
    if args is None:
        m = state["cfg"]["model"]   # a Namespace for legacy (non-dataclass) models such as diff_discrete, else a dict / DictConfig
        args = argparse.Namespace(**(vars(m) if isinstance(m, argparse.Namespace) else dict(m)))
    if not hasattr(args, "task"):
        args.task = "speech_diffusion_discrete"
    if getattr(args, "arch", None) not in ("diff_discrete",):
        args.arch = "diff_discrete"
    args.speech_decoder_ckpt = None  # the VAE weights travel inside the diffusion checkpoint (encoder.speech_decoder.*)
    task = compat.setup_task(args)
    model = task.build_model(args, from_checkpoint=True)
    model.load_state_dict(state["model"], strict=True)
    return model.to(device).eval(), task


def write_shards(lines, output_dir: str, split: str, rank: int, world: int) -> str:
    """The only cross-rank step of the normalization path, and it is host-side: every rank writes its own restartable shard
    `{split}.rank{r}.tsv` ({item index: TSV line} of the utterances plan_batches gave it); rank 0 gathers the dictionaries (gloo,
    objects — no tensor ever crosses ranks) and writes `{split}.tsv` in the original utterance order, i.e. the file a single
    process writes (diff_norm_synthesis.py:218-222).  Returns the path this rank wrote last."""
    path = os.path.join(output_dir, f"{split}.tsv" if world == 1 else f"{split}.rank{rank}.tsv")
    write_tsv(path, [lines[i] for i in sorted(lines)])
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("gloo")
        gathered = [None] * world
        dist.all_gather_object(gathered, lines)
        if rank == 0:
            merged = {}
            for g in gathered:
                if set(g) & set(merged):
                    raise RuntimeError("two ranks normalized the same utterance: the sharding plan is not a partition")
                merged.update(g)
            path = os.path.join(output_dir, f"{split}.tsv")
            write_tsv(path, [merged[i] for i in sorted(merged)])
    return path


def main(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("diffnorm_b200 needs a CUDA device (no CPU path)")
    torch.cuda.set_device(local)
    os.makedirs(args.output_dir, exist_ok=True)
    model, _ = load_model(args.model_ckpt, f"cuda:{local}")
    eng = model.encoder._engine()
    eng.reserve(args.max_tokens)
    runner = NormalizationRunner(eng, start_step=args.start_step, max_tokens=args.max_tokens)
    for split in args.splits.split(","):
        items, unfound = prepare_data(args.reduce_tsv_dir, args.orig_tsv_dir, args.feature_dir, split)
        print("Unfound: ", unfound)
        lines = runner.run_items(items, rank=rank, world_size=world)
        write_shards(lines, args.output_dir, split, rank, world)
        print("Finished processing ", split)


def cli_main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--reduce_tsv_dir", type=str, default=None, help="path to reduced tsv dir")
    p.add_argument("--orig_tsv_dir", type=str, default=None, help="path to original tsv dir")
    p.add_argument("--dummy-config", type=str, default=None, help="path to a dummy config file (unused)")
    p.add_argument("--feature_dir", type=str, default=None, help="path to target vae feats")
    p.add_argument("--model_ckpt", type=str, default=None, help="path to the diffusion model checkpoint")
    p.add_argument("--start_step", type=int, default=50)
    p.add_argument("--output_dir", type=str, help="path to output dir")
    p.add_argument("--splits", type=str, default="test,dev,train")
    p.add_argument("--max-tokens", type=int, default=64000, help="padded frames per batch (length-bucketed)")
    main(p.parse_args(argv))


if __name__ == "__main__":
    cli_main(sys.argv[1:])
