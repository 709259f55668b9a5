"""The caller path on the GPU: TSV + .npy inputs -> device reduce / gather-pad ->
pass -> device reduce -> TSV lines, against the oracle run on the same gathered features and the same noise."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from diffnorm_b200 import data  # noqa: E402
from diffnorm_b200.engine import DiffNormEngine  # noqa: E402
from oracle import diffnorm_oracle as O  # noqa: E402


def test_runner_end_to_end_against_oracle(tmp_path):
    rng = np.random.default_rng(3)
    red, orig, feat_dir = tmp_path / "red", tmp_path / "orig", tmp_path / "feat" / "test"
    for d in (red, orig, feat_dir):
        d.mkdir(parents=True)
    rows_o, rows_r, feats, units_full = [], [], {}, {}
    for k, n_runs in enumerate([37, 12, 50, 1, 23]):
        ids = rng.integers(0, 1000, size=n_runs)
        ids[1:][ids[1:] == ids[:-1]] += 1                      # adjacent runs differ
        durs = rng.geometric(0.6, size=n_runs)
        full = np.repeat(ids, durs)
        f = rng.standard_normal((len(full), 768)).astype(np.float32)
        uid = f"utt{k}"
        np.save(feat_dir / f"{uid}.feat.npy", f)
        feats[uid], units_full[uid] = f, full
        rows_o.append(f"{uid}\t{uid}.wav\t{100 + k}\t{' '.join(map(str, full))}\t{len(full)}")
        rows_r.append(f"{uid}\t{uid}.wav\t{100 + k}\t{' '.join(map(str, ids))}\t{n_runs}")
    (orig / "test.tsv").write_text(data.HEADER + "\n" + "\n".join(rows_o) + "\n")
    (red / "test.tsv").write_text(data.HEADER + "\n" + "\n".join(rows_r) + "\n")
    items, unfound = data.prepare_data(str(red), str(orig), str(tmp_path / "feat"), "test")
    assert unfound == 0 and len(items) == 5

    arch = O.Arch(latent_dim=16)
    sd = O.init_state_dict(arch, seed=1, gains=O.PARITY_GAINS)
    eng = DiffNormEngine(sd, "cuda")
    start = 5
    runner = data.NormalizationRunner(eng, start_step=start, max_tokens=4096)

    torch.manual_seed(99)
    lines = runner.run_items(items)
    assert sorted(lines) == [0, 1, 2, 3, 4]

    # replay: one batch holding all five utterances in length order, same device RNG stream
    plan = data.plan_batches([it.reduce_tgt_n_frames for it in items], 4096)[0]
    assert len(plan) == 1
    idx = plan[0]
    torch.manual_seed(99)
    res, dev_units, feat_dev = runner.normalize_batch([feats[items[i].audio_id] for i in idx],
                                                      [units_full[items[i].audio_id] for i in idx], return_units=True)
    B, T = len(idx), max(items[i].reduce_tgt_n_frames for i in idx)
    # the engine draws its noise with the library's Philox kernel, seeded from torch's CPU generator: replay both draws
    from diffnorm_b200 import ops
    torch.manual_seed(99)
    eps_vae = ops.randn((B, 16, T), "cuda", eng._noise_seed(), 0).cpu()
    eps_q = ops.randn((B, T, 16), "cuda", eng._noise_seed(), 0).cpu()
    assert abs(float(eps_q.mean())) < 0.05 and abs(float(eps_q.std()) - 1.0) < 0.05
    # oracle pre-processing: reduce_token(full) -> index_to_keep -> gather -> pad (diff_norm_synthesis.py:150-169)
    lens = torch.tensor([items[i].reduce_tgt_n_frames for i in idx])
    ofeat = torch.zeros(B, T, 768)
    for j, i in enumerate(idx):
        uid = items[i].audio_id
        dd, du, keep = O.reduce_tgt(units_full[uid].tolist())
        assert len(keep) == items[i].reduce_tgt_n_frames
        ofeat[j, : len(keep)] = torch.from_numpy(feats[uid][keep])
    assert torch.equal(feat_dev, ofeat)                          # device reduce + gather + pad is bit-exact
    mask = O.lengths_to_mask(lens, T)
    ref = O.normalize_pass(sd, arch, ofeat, mask, start, eps_vae, eps_q)
    top2 = ref["logits"].topk(2, dim=-1).values
    conf = (top2[..., 0] - top2[..., 1]) > 0.1 * float(ref["logits"].std())
    agree = tot = 0
    for j, i in enumerate(idx):
        n = int(lens[j])
        a = torch.from_numpy(dev_units[j]) == ref["out_tokens"][j]
        agree += int(a[conf[j, :n]].sum())
        tot += int(conf[j, :n].sum())
        # TSV line = reduce of the device's own units, n_frames = length before the second reduce (:211-222)
        dd, _, _ = O.reduce_tgt(dev_units[j].tolist())
        it = items[i]
        assert lines[int(i)] == f"{it.audio_id}\t{it.src_audio}\t{it.src_n_frames}\t{' '.join(map(str, dd))}\t{n}"
        assert res[j][0].tolist() == dd and res[j][1] == n
    print(f"[parity] runner: unit agreement on confident frames {agree}/{tot}")
    assert agree >= 0.995 * tot
    # length mismatch is an error, like the reference's assert (:152)
    with pytest.raises(AssertionError):
        runner.normalize_batch([feats["utt0"]], [units_full["utt0"]], expect_reduced=[999])


def test_quantize_cli_end_to_end(tmp_path):
    """quantize_with_kmeans.py drop-in: centroids + feature manifest -> '{name}|{units}' lines, labels equal
    to the float64 oracle's nearest centroid."""
    from diffnorm_b200 import quantize_cli
    rng = np.random.default_rng(2)
    centers = rng.standard_normal((1000, 768)).astype(np.float32)
    np.save(tmp_path / "centers.npy", centers)
    fdir = tmp_path / "feat"
    fdir.mkdir()
    feats = {}
    rows = [str(fdir)]
    for k, n in enumerate([17, 1, 230]):
        lab = rng.integers(0, 1000, size=n)
        f = (centers[lab] + 0.5 * rng.standard_normal((n, 768))).astype(np.float32)
        np.save(fdir / f"u{k}.feat.npy", f)
        feats[f"u{k}.feat.npy"] = f
        rows.append(f"u{k}.feat.npy\t{n}")
    (tmp_path / "split.manifest.tsv").write_text("\n".join(rows) + "\n")
    out = tmp_path / "out" / "split.quant.tsv"
    quantize_cli.cli_main(["--kmeans_model_path", str(tmp_path / "centers.npy"), "--manifest_path", str(tmp_path / "split.manifest.tsv"),
                           "--out_quantized_file_path", str(out)])
    lines = out.read_text().strip().splitlines()
    assert len(lines) == 3
    for ln in lines:
        name, units = ln.split("|")
        want = O.kmeans_predict(centers, feats[name])
        assert [int(u) for u in units.split(" ")] == want.tolist()
