"""TEST INFRASTRUCTURE ONLY.  A tiny synthetic corpus in the on-disk formats around the path: source
and target feature manifests + `{id}.feat.npy` files + the unit TSV.  Deterministic from the seed, so the golden vectors
minted from the reference's dataset class (oracle/make_golden.py make_dataset) and the tests see the same files.
Covers: ragged lengths, long runs, an utterance missing from the target manifest, one whose unit count differs from its
feature length, blank lines, and a unit id outside the dictionary (-> <unk>)."""
import os

import numpy as np

HEADER = "id\tsrc_audio\tsrc_n_frames\ttgt_audio\ttgt_n_frames"


def write_corpus(root: str, split: str = "train", n: int = 9, seed: int = 5, dim: int = 32):
    rng = np.random.default_rng(seed)
    src_dir, tgt_dir = os.path.join(root, "src_feat"), os.path.join(root, "tgt_feat")
    for d in (src_dir, tgt_dir, os.path.join(root, "tsv")):
        os.makedirs(os.path.join(d, split) if d != os.path.join(root, "tsv") else d, exist_ok=True)
    src_rows, tgt_rows, tsv_rows = [], [], []
    for i in range(n):
        uid = f"utt_{i:03d}"
        n_src, n_tgt = int(rng.integers(5, 40)), int(rng.integers(6, 60))
        units = np.repeat(rng.integers(0, 1000, size=n_tgt), rng.integers(1, 4, size=n_tgt))[:n_tgt]
        if i == 2:
            units[3] = 1000      # outside the 1000-unit dictionary -> <unk>
        np.save(os.path.join(src_dir, split, f"{uid}.feat.npy"), rng.standard_normal((n_src, dim)).astype(np.float32))
        np.save(os.path.join(tgt_dir, split, f"{uid}.feat.npy"), rng.standard_normal((n_tgt, dim)).astype(np.float32))
        src_rows.append(f"{uid}.feat.npy\t{n_src}")
        if i != 4:               # utt_004 is missing from the target manifest -> skipped
            tgt_rows.append(f"{uid}.feat.npy\t{n_tgt if i != 6 else n_tgt + 1}")   # utt_006: length mismatch -> skipped
        tsv_rows.append(f"{uid}\t{uid}.wav\t{n_src * 320}\t{' '.join(str(int(u)) for u in units)}\t{n_tgt}")
    with open(os.path.join(src_dir, f"{split}.manifest.tsv"), "w") as f:
        f.write(os.path.join(src_dir, split) + "\n" + "\n".join(src_rows) + "\n\n")
    with open(os.path.join(tgt_dir, f"{split}.manifest.tsv"), "w") as f:
        f.write(os.path.join(tgt_dir, split) + "\n" + "\n".join(tgt_rows) + "\n")
    with open(os.path.join(root, "tsv", f"{split}.tsv"), "w") as f:
        f.write(HEADER + "\n" + "\n".join(tsv_rows[:5]) + "\n\n" + "\n".join(tsv_rows[5:]) + "\n")
    return src_dir, tgt_dir, os.path.join(root, "tsv")
