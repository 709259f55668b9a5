"""The training dataset + collater behind `task.load_dataset` (repr_to_repr_unit_dataset.py:46-399) against golden
batches minted from the reference's OWN class and Dictionary on the synthetic corpus of oracle/dataset_fixture.py
(oracle/make_golden.py make_dataset), and, where /root/reference exists, against the live class field by field."""
import argparse
import os

import numpy as np
import pytest
import torch

from diffnorm_b200 import data
from diffnorm_b200.plugin import compat
from oracle import ref_loader
from oracle.dataset_fixture import write_corpus

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("id", "target", "target_unit", "reduce_target", "reduce_target_unit", "target_lengths", "reduce_target_lengths")


def _task(root, src_dir, tgt_dir, tsv_dir, name="speech_diffusion_discrete"):
    args = argparse.Namespace(task=name, data=tsv_dir, src_feat_dir=src_dir, tgt_feat_dir=tgt_dir, target_is_code=True,
                              target_code_size=1000, dummy_config=None, seed=1)
    return compat.setup_task(args)


def test_load_dataset_and_collate_match_reference_golden(tmp_path):
    g = np.load(os.path.join(GOLD, "dataset_collate.npz"))
    src_dir, tgt_dir, tsv_dir = write_corpus(str(tmp_path))
    task = _task(tmp_path, src_dir, tgt_dir, tsv_dir)
    task.load_dataset("train")
    ds = task.dataset("train")
    assert ds.ids == g["ids"].tolist()                       # utt_004 (no target manifest row) and utt_006 (length) skipped
    assert np.array_equal(ds.ordered_indices(), g["ordered_indices"]) and np.array_equal(ds.sizes, g["sizes"])
    for b in range(2):
        batch = ds.collater([ds[int(i)] for i in g[f"b{b}_idx"]])
        for k in FIELDS:
            got = batch[k].numpy()
            assert got.dtype == g[f"b{b}_{k}"].dtype and np.array_equal(got, g[f"b{b}_{k}"]), (b, k)
        assert np.array_equal(batch["net_input"]["src_tokens"].numpy(), g[f"b{b}_src_tokens"])
        assert np.array_equal(batch["net_input"]["src_lengths"].numpy(), g[f"b{b}_src_lengths"])
        assert batch["ntokens"] == int(g[f"b{b}_ntokens"]) and batch["nsentences"] == int(g[f"b{b}_nsentences"])
        assert batch["net_input"]["prev_output_tokens"] is None and batch["speaker"] is None
    # the schema the criterions rely on: 0-padding, unit k -> k + 4, out-of-dictionary unit -> <unk> (3)
    item = ds[2]
    assert int(item.tgt_unit[3]) == 3 and int(item.tgt_unit.min()) >= 3
    units = [int(u) for u in open(os.path.join(tsv_dir, "train.tsv")).read().splitlines()[1].split("\t")[3].split(" ")]
    assert ds[0].tgt_unit.tolist() == [u + 4 for u in units]
    assert ds.collater([]) == {}


def test_reduce_tgt_edge_cases_and_batch_sampler(tmp_path):
    R = data.ReprToReprUnitDataset._reduce_tgt
    assert R([])[0] == [] and R([])[1] == [1] and R([])[2].numel() == 0      # the reference quirk (:112)
    d, du, k = R([7])
    assert (d, du, k.tolist()) == ([7], [1], [0])
    d, du, k = R([5, 5, 6, 6, 6, 7, 5])
    assert (d, du, k.tolist()) == ([5, 6, 7, 5], [2, 3, 1, 1], [0, 2, 5, 6]) and k.dtype == torch.long
    from oracle import diffnorm_oracle as O
    rng = np.random.default_rng(0)
    for _ in range(50):
        t = np.repeat(rng.integers(0, 5, size=30), rng.integers(1, 4, size=30)).tolist()
        d, du, k = R(t)
        assert (d, du, k.tolist()) == O.reduce_tgt(t)
    src_dir, tgt_dir, tsv_dir = write_corpus(str(tmp_path), n=40, seed=9)
    ds = _task(tmp_path, src_dir, tgt_dir, tsv_dir, "speech_decoder")
    ds.load_dataset("train")
    ds = ds.dataset("train")
    batches = ds.batch_sampler(max_tokens=120)
    assert sorted(np.concatenate(batches).tolist()) == list(range(len(ds)))
    assert all(len(b) * ds.sizes[b].max() <= 120 for b in batches)


def test_eval_split_cap_and_shuffle_flag(tmp_path):
    src_dir, tgt_dir, tsv_dir = write_corpus(str(tmp_path), split="dev", n=5, seed=2)
    (tmp_path / "cfg.yaml").write_text("shuffle: true\n")
    args = argparse.Namespace(data=tsv_dir, src_feat_dir=src_dir, tgt_feat_dir=tgt_dir, dummy_config=str(tmp_path / "cfg.yaml"))
    d = compat.Dictionary()
    ds = data.ReprToReprUnitDataset.from_manifest(args, "dev", d)
    assert ds.shuffle is False                         # shuffle only applies to train splits (:75)
    write_corpus(str(tmp_path), split="train", n=5, seed=2)
    assert data.ReprToReprUnitDataset.from_manifest(args, "train", d).shuffle is True


@pytest.mark.ref
@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference")
def test_against_live_reference_class(tmp_path):
    ns = ref_loader.load_dataset_module()
    rd = ns.Dictionary()
    for i in range(1000):
        rd.add_symbol(str(i))
    src_dir, tgt_dir, tsv_dir = write_corpus(str(tmp_path), n=14, seed=21)
    ref = ns.module.ReprToReprUnitDatasetCreator.from_tsv(src_dir, tgt_dir, tsv_dir, ns.S2SDataConfig(False), "train", True,
                                                          1, 1, tgt_dict=rd)
    task = _task(tmp_path, src_dir, tgt_dir, tsv_dir)
    task.load_dataset("train")
    ds = task.dataset("train")
    assert len(ds) == len(ref) and ds.ids == ref.ids
    assert np.array_equal(ds.ordered_indices(), ref.ordered_indices())
    order = ds.ordered_indices().tolist()
    for idx in (order[:5], order[5:]):
        a, b = ds.collater([ds[i] for i in idx]), ref.collater([ref[i] for i in idx])
        assert set(a) == set(b)
        for k in FIELDS:
            assert torch.equal(a[k], b[k]) and a[k].dtype == b[k].dtype, k
        assert torch.equal(a["net_input"]["src_tokens"], b["net_input"]["src_tokens"])
        assert torch.equal(a["net_input"]["src_lengths"], b["net_input"]["src_lengths"])
        assert a["ntokens"] == b["ntokens"] and a["nsentences"] == b["nsentences"]
        for i in idx:
            x, y = ds._reduce_tgt(ds.tgt_units[i]), ref._reduce_tgt(ref.tgt_units[i])
            assert x[0] == y[0] and x[1] == y[1] and torch.equal(x[2], y[2])
