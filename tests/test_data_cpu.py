"""Host-side caller logic: TSV/manifest readers, the native batcher against golden vectors from the
reference's compiled Cython (oracle/_ref, built from /root/reference/fairseq/data/data_utils_fast.pyx in the authoring
container by oracle/Makefile), sharding plan properties.  CPU only."""
import os

import numpy as np
import pytest

from diffnorm_b200 import data

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_batcher_matches_reference_cython_golden():
    g = np.load(os.path.join(GOLD, "batch_by_size.npz"))
    n_cases = int(g["n_cases"])
    assert n_cases >= 200
    for i in range(n_cases):
        toks = g[f"toks_{i}"]
        mt, ms, bm = (int(v) for v in g[f"args_{i}"])
        got = [e - s for s, e in data.batch_by_size(toks, mt, ms, bm)]
        assert got == g[f"sizes_{i}"].tolist(), (i, toks.tolist(), mt, ms, bm)


def test_batcher_properties_and_errors():
    rng = np.random.default_rng(1)
    lens = np.sort(rng.integers(200, 2001, size=5000))
    ranges = data.batch_by_size(lens, max_tokens=64000)
    assert ranges[0][0] == 0 and ranges[-1][1] == len(lens)
    for (s, e), (s2, _) in zip(ranges, ranges[1:]):
        assert e == s2
    for s, e in ranges:
        assert (e - s) * lens[s:e].max() <= 64000
    assert data.batch_by_size([], 100) == []
    with pytest.raises(Exception):
        data.batch_by_size([10, 500], max_tokens=100)


def test_plan_batches_covers_everything_and_balances():
    rng = np.random.default_rng(2)
    lens = np.clip(np.round(np.exp(rng.normal(np.log(600), 0.5, size=20000))), 200, 2000).astype(np.int64)
    for world in (1, 2, 8):
        plan = data.plan_batches(lens, max_tokens=64000, world_size=world)
        seen = np.concatenate([np.concatenate(p) for p in plan])
        assert sorted(seen.tolist()) == list(range(len(lens)))          # every utterance exactly once
        loads = [sum(len(b) * float(data.pass_cost(lens[b].max())) for b in p) for p in plan]
        assert max(loads) / (sum(loads) / world) < 1.05                  # LPT keeps ranks within 5 %
        for p in plan:
            for b in p:
                assert len(b) * lens[b].max() <= 64000


def test_tsv_and_manifest_readers(tmp_path):
    red, orig, feat = tmp_path / "red", tmp_path / "orig", tmp_path / "feat" / "dev"
    for d in (red, orig, feat):
        d.mkdir(parents=True)
    (orig / "dev.tsv").write_text(data.HEADER + "\nu1\ta.wav\t30\t5 5 6 6 6 7\t6\nbad line\nu2\tb.wav\t10\t1 1\t2\nu3\tc.wav\t5\t9\t1\n")
    (red / "dev.tsv").write_text(data.HEADER + "\nu1\ta.wav\t30\t5 6 7\t3\nu2\tb.wav\t10\t1\t1\n")
    np.save(feat / "u1.feat.npy", np.zeros((6, 768), np.float32))
    items, unfound = data.prepare_data(str(red), str(orig), str(tmp_path / "feat"), "dev")
    assert [it.audio_id for it in items] == ["u1"] and unfound == 2  # u2 has no feature file, u3 no reduced row
    it = items[0]
    assert (it.tgt_unit, it.tgt_n_frames, it.reduce_tgt_unit, it.reduce_tgt_n_frames) == ("5 5 6 6 6 7", 6, "5 6 7", 3)
    (tmp_path / "dev.manifest.tsv").write_text(f"{feat}\nu1.feat.npy\t6\n")
    root, rows = data.read_manifest(str(tmp_path / "dev.manifest.tsv"))
    assert root == str(feat) and rows == [("u1.feat.npy", 6)]
    data.write_tsv(str(tmp_path / "out.tsv"), ["u1\ta.wav\t30\t5 6\t3"])
    assert (tmp_path / "out.tsv").read_text().splitlines() == [data.HEADER, "u1\ta.wav\t30\t5 6\t3"]
