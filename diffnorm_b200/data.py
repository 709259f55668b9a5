"""The caller side of the normalization pass: manifest / unit-TSV reader, native length-bucketed
batching, utterance sharding across GPUs, the batched GPU pre/post-processing around ``ddim_sample`` and the
output TSV writer.  Mirrors research/TranSpeech/diff_norm_synthesis.py:70-222 (file formats: SURVEY Appendix B)
but replaces its fixed 100-utterance file-order batches and per-utterance Python loops:

  reference (per utterance, Python)                     here (per batch, device)
  reduce_token(full_unit) -> index_to_keep   :150       dn_reduce_tgt on the padded original units
  tgt_feat[index_to_keep]; zero pad          :151-169   dn_gather_pack from the packed feature rows
  ddim_sample                                :204       DiffNormEngine.normalize
  .cpu().tolist(); reduce_token again        :213-216   dn_reduce_tgt on the predicted units, one D2H per batch
"""
from __future__ import annotations

import ctypes
import os
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

HEADER = "id\tsrc_audio\tsrc_n_frames\ttgt_audio\ttgt_n_frames"


@dataclass
class UtteranceItem:  # diff_norm_synthesis.py:57-67 AllDataItem
    audio_id: str
    src_audio: str
    src_n_frames: int
    tgt_unit: str
    tgt_n_frames: int
    reduce_tgt_unit: str
    reduce_tgt_n_frames: int
    feature_file: str


def read_unit_tsv(path: str) -> "OrderedDict[str, Tuple[str, int, str, int]]":
    """Header line skipped; rows with != 5 tab-separated fields skipped (diff_norm_synthesis.py:79-88)."""
    out: "OrderedDict[str, Tuple[str, int, str, int]]" = OrderedDict()
    with open(path, "r") as f:
        f.readline()
        for line in f:
            parts = line.strip().split("\t")
            if len(parts) != 5:
                continue
            audio_id, src_audio, src_n, tgt_audio, tgt_n = parts
            out[audio_id] = (src_audio, int(src_n), tgt_audio, int(tgt_n))
    return out


def prepare_data(reduce_tsv_dir: str, orig_tsv_dir: str, feature_dir: str, split: str) -> Tuple[List[UtteranceItem], int]:
    """Join the reduced and original unit TSVs and keep utterances whose ``{id}.feat.npy`` exists
    (diff_norm_synthesis.py:70-116).  Order = order of the original TSV.  Returns (items, unfound)."""
    reduce_map = read_unit_tsv(os.path.join(reduce_tsv_dir, f"{split}.tsv"))
    items: "OrderedDict[str, UtteranceItem]" = OrderedDict()
    unfound = 0
    for audio_id, (_, _, tgt_audio, tgt_n) in read_unit_tsv(os.path.join(orig_tsv_dir, f"{split}.tsv")).items():
        feat = os.path.join(feature_dir, split, f"{audio_id}.feat.npy")
        if audio_id not in reduce_map or not os.path.exists(feat):
            unfound += 1
            continue
        r_src, r_src_n, r_units, r_n = reduce_map[audio_id]
        items[audio_id] = UtteranceItem(audio_id, r_src, r_src_n, tgt_audio, tgt_n, r_units, r_n, feat)
    return list(items.values()), unfound


def read_manifest(path: str) -> Tuple[str, List[Tuple[str, int]]]:
    """``{split}.manifest.tsv``: line 1 = feature dir, then ``{id}.feat.npy \\t N_full``
    (speech2unit/pretrained/utils.py:131-141, repr_to_repr_unit_dataset.py:311-323)."""
    with open(path) as f:
        root = f.readline().strip()
        rows = []
        for line in f:
            parts = line.strip().split("\t")
            if len(parts) == 2:
                rows.append((parts[0], int(parts[1])))
    return root, rows


# ------------------------------------------------------------------------------------------------ batching
def batch_by_size(num_tokens: Sequence[int], max_tokens: int = 0, max_sentences: int = 0, bsz_mult: int = 1) -> List[Tuple[int, int]]:
    """Native (C++) equivalent of fairseq's ``batch_by_size_vec`` (data_utils_fast.pyx:20-101): half-open index
    ranges over ``num_tokens`` in the given order."""
    from ._lib import DiffNormLibraryError, lib
    toks = np.ascontiguousarray(num_tokens, dtype=np.int64)
    n = len(toks)
    ends = np.zeros(n + 1, dtype=np.int64)
    k = lib.dn_batch_by_size(toks.ctypes.data_as(ctypes.c_void_p), n, int(max_tokens), int(max_sentences), int(bsz_mult),
                             ends.ctypes.data_as(ctypes.c_void_p))
    if k < 0:
        raise DiffNormLibraryError(f"dn_batch_by_size failed ({k}): an utterance exceeds max_tokens={max_tokens}?")
    starts = np.concatenate([[0], ends[: k - 1]]) if k > 0 else np.zeros(0, dtype=np.int64)
    return [(int(s), int(e)) for s, e in zip(starts, ends[:k])]


def pass_cost(n: np.ndarray) -> np.ndarray:
    """Relative device cost of normalizing an utterance of n frames: GEMM/conv work is linear in n, attention is
    quadratic (D(z, N) = 141.78 M + 12,288 N MAC per frame)."""
    n = np.asarray(n, dtype=np.float64)
    return n * (141.78e6 + 12288.0 * n)


def plan_batches(lengths: Sequence[int], max_tokens: int, max_sentences: int = 0, bsz_mult: int = 1, world_size: int = 1,
                 pad_multiple: int = 1) -> List[List[np.ndarray]]:
    """Sort utterances by length, cut them into batches under a padded-token budget, and assign the batches to
    `world_size` ranks by longest-processing-time-first on the cost model (no collective is ever needed:
    utterances are independent).  Returns, per rank, a list of index arrays (indices into `lengths`)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(lengths, kind="stable")
    padded = (lengths[order] + pad_multiple - 1) // pad_multiple * pad_multiple
    ranges = batch_by_size(padded, max_tokens, max_sentences, bsz_mult)
    batches = [order[s:e] for s, e in ranges]
    costs = np.array([len(b) * float(pass_cost(padded[s:e].max())) for b, (s, e) in zip(batches, ranges)])
    load = np.zeros(world_size)
    per_rank: List[List[np.ndarray]] = [[] for _ in range(world_size)]
    for bi in np.argsort(-costs, kind="stable"):
        r = int(np.argmin(load))
        per_rank[r].append(batches[bi])
        load[r] += costs[bi]
    return per_rank


# ------------------------------------------------------------------------------------------------ runner
class NormalizationRunner:
    """Drives DiffNormEngine over lists of utterances (features + original unit strings) and returns TSV rows.

    A batch goes through three stages so that disk, host and device overlap (SURVEY §8f-1 "overlap I/O with compute"):
      stage    host only: np.load, pack the feature rows and pad the unit rows into pinned buffers  (prefetch thread)
      launch   enqueue H2D + reduce + gather + the whole pass; no host synchronisation when the reduced lengths are known
               from the TSV (the padded length T then comes from the manifest, not from the device)
      collect  ONE device-to-host read per batch (reduced units + both count vectors), the reference's length assert
    `run_items` keeps one batch staged ahead and one launched ahead of the one being collected."""

    def __init__(self, engine, start_step: int = 50, max_tokens: int = 64000, max_sentences: int = 0, sampler: str = "ddim"):
        self.eng, self.start_step, self.max_tokens, self.max_sentences, self.sampler = engine, start_step, max_tokens, max_sentences, sampler

    @staticmethod
    def stage(feats: List[np.ndarray], full_units: List[np.ndarray]):
        import torch
        n_full = np.array([len(u) for u in full_units], dtype=np.int64)
        for f, u in zip(feats, full_units):
            if f.shape[0] != len(u):
                raise ValueError(f"feature rows {f.shape[0]} != number of original units {len(u)}")
        B, Tf = len(feats), int(n_full.max())
        units_h = torch.zeros(B, Tf, dtype=torch.int64).pin_memory()
        for i, u in enumerate(full_units):
            units_h[i, : len(u)] = torch.from_numpy(np.asarray(u, dtype=np.int64))
        packed_h = torch.from_numpy(np.concatenate([np.asarray(f, dtype=np.float32) for f in feats], axis=0)).pin_memory()
        row0_h = torch.from_numpy(np.concatenate([[0], np.cumsum(n_full)[:-1]]).astype(np.int64)).pin_memory()
        lens_h = torch.from_numpy(n_full.astype(np.int32)).pin_memory()
        return dict(B=B, units=units_h, packed=packed_h, row0=row0_h, lens=lens_h)

    def launch(self, st, expect_reduced: Optional[Sequence[int]] = None):
        from . import ops
        dev = self.eng.dev
        units_d = st["units"].to(dev, non_blocking=True)
        packed = st["packed"].to(dev, non_blocking=True)
        lens_full = st["lens"].to(dev, non_blocking=True)
        _, _, keep, counts = ops.reduce_tgt(units_d, lens_full)   # first reduce: index_to_keep of the ORIGINAL units (:150)
        if expect_reduced is not None:
            T = int(max(expect_reduced))                          # known from the TSV: nothing to wait for
        else:
            T = int(counts.max().item())
        T = max(1, min(T, keep.shape[1]))
        keep_t = keep[:, :T].contiguous()
        feat = ops.gather_pack(packed, st["row0"].to(dev, non_blocking=True), keep_t, counts, T)  # [B,T,768], zero padded (:164-169)
        out = self.eng.normalize(feat, counts, self.start_step, sampler=self.sampler)
        return dict(B=st["B"], counts=counts, out=out, feat=feat, expect=None if expect_reduced is None else list(expect_reduced))

    @staticmethod
    def collect(h, return_units: bool = False):
        import torch
        out = h["out"]
        counts_h, dedup, cnt2 = (t.to("cpu", non_blocking=True) for t in (h["counts"], out["dedup"], out["counts"]))
        units = out["units"].to("cpu", non_blocking=True) if return_units else None
        torch.cuda.synchronize()
        counts_h, dedup, cnt2 = counts_h.numpy(), dedup.numpy(), cnt2.numpy()
        if h["expect"] is not None and list(counts_h) != h["expect"]:   # the reference's assert (:152)
            raise AssertionError("reduced length from the original units does not match reduce_tgt_n_frames")
        res = [(dedup[i, : cnt2[i]].copy(), int(counts_h[i])) for i in range(h["B"])]
        if return_units:
            units = units.numpy()
            return res, [units[i, : counts_h[i]].copy() for i in range(h["B"])], h["feat"].cpu()
        return res

    def normalize_batch(self, feats: List[np.ndarray], full_units: List[np.ndarray], expect_reduced: Optional[List[int]] = None,
                        return_units: bool = False):
        """feats[i] fp32 [N_full_i, 768], full_units[i] int64 [N_full_i] -> per utterance (reduced units, n_frames)
        where n_frames = frames fed to the model = length BEFORE the second reduce (diff_norm_synthesis.py:211-222)."""
        return self.collect(self.launch(self.stage(feats, full_units), expect_reduced), return_units)

    def run_items(self, items: List[UtteranceItem], rank: int = 0, world_size: int = 1, progress=None) -> Dict[int, str]:
        """Normalizes this rank's share of `items`; returns {item index: TSV line}."""
        from concurrent.futures import ThreadPoolExecutor
        lengths = [it.reduce_tgt_n_frames for it in items]
        plan = plan_batches(lengths, self.max_tokens, self.max_sentences, world_size=world_size)[rank]
        lines: Dict[int, str] = {}

        def load(idx):
            feats = [np.load(items[i].feature_file) for i in idx]
            units = [np.array([int(x) for x in items[i].tgt_unit.split(" ")], dtype=np.int64) for i in idx]
            return self.stage(feats, units)

        def finish(idx, handle):
            for i, (red, n_frames) in zip(idx, self.collect(handle)):
                it = items[i]
                lines[int(i)] = f"{it.audio_id}\t{it.src_audio}\t{it.src_n_frames}\t{' '.join(str(int(v)) for v in red)}\t{n_frames}"
            if progress is not None:
                progress(len(idx))

        with ThreadPoolExecutor(max_workers=1) as pool:
            staged = pool.submit(load, plan[0]) if plan else None
            in_flight = None
            for k, idx in enumerate(plan):
                st = staged.result()
                staged = pool.submit(load, plan[k + 1]) if k + 1 < len(plan) else None     # disk + pinning under the GPU's work
                handle = self.launch(st, [items[i].reduce_tgt_n_frames for i in idx])
                if in_flight is not None:
                    finish(*in_flight)                                                       # batch k-1, while batch k runs
                in_flight = (idx, handle)
            if in_flight is not None:
                finish(*in_flight)
        return lines


# ------------------------------------------------------------------------------------------------ training dataset
@dataclass
class ReprToReprDatasetItem:  # repr_to_repr_unit_dataset.py:34-42
    index: int
    src_feat: "torch.Tensor"
    tgt_feat: "torch.Tensor"
    tgt_unit: "torch.Tensor"
    reduce_tgt_unit: "torch.Tensor"
    reduce_tgt_feat: "torch.Tensor"


def _feat_manifest(path: str) -> Dict[str, Tuple[str, str]]:
    """repr_to_repr_unit_dataset.py:311-323: id (file name up to the first dot) -> (feature path, length string)."""
    root, rows = None, {}
    with open(path, "r") as f:
        root = f.readline().strip()
        for line in f:
            if len(line.strip()) == 0:
                continue
            name, n = line.strip().split("\t")
            rows[name.split(".")[0]] = (f"{root}/{name}", n)
    return rows


def load_samples_from_tsv(src_feat_dir: str, tgt_feat_dir: str, raw_audio_root: str, split: str, log=print) -> List[Dict]:
    """ReprToReprUnitDatasetCreator._load_samples_from_tsv (repr_to_repr_unit_dataset.py:309-369): join the source and
    target feature manifests with the unit TSV; utterances missing from either manifest or whose unit count differs from
    the target feature length are skipped; non-train splits stop after the 4001st kept utterance (:365-368)."""
    src = _feat_manifest(f"{src_feat_dir}/{split}.manifest.tsv")
    tgt = _feat_manifest(f"{tgt_feat_dir}/{split}.manifest.tsv")
    samples: List[Dict] = []
    with open(f"{raw_audio_root}/{split}.tsv") as f:
        f.readline()
        for line in f:
            if len(line.strip()) == 0:
                continue
            uid, _src_audio, _src_n, units, _tgt_n = line.rstrip().split("\t")
            if uid not in src or uid not in tgt:
                log(f"src_id: {uid} not found in feat manifest")
                continue
            tokens = [int(x) for x in units.split(" ")]
            if len(tokens) != int(tgt[uid][1]):
                log(f"warning: mismatched feature and unit size. tgt_tokens: {len(tokens)}, tgt_feat_len: {tgt[uid][1]}")
                continue
            samples.append({"id": uid, "src_audio": src[uid][0], "src_n_frames": src[uid][1], "tgt_audio": tgt[uid][0],
                            "tgt_unit": tokens, "tgt_n_frames": tgt[uid][1]})
            if "train" not in split and len(samples) > 4000:
                break
    return samples


def _dataset_base():
    try:  # pragma: no cover - only where fairseq is installed
        from fairseq.data import FairseqDataset  # type: ignore
        return FairseqDataset
    except Exception:  # noqa: BLE001
        import torch.utils.data
        return torch.utils.data.Dataset


class ReprToReprUnitDataset(_dataset_base()):
    """Training / validation dataset of both tasks (repr_to_repr_unit_dataset.py:46-258): per utterance the source
    features, the target mHuBERT features, the target units, and their run-length-reduced forms; `collater` builds the
    sample dict the two criterions read: 0-padding, unit k -> dictionary index k + 4, `ntokens` = sum
    of the reduced lengths, rows sorted by source length (descending)."""

    def __init__(self, split: str, is_train_split: bool, audio_paths: List[str], tgt_feat_paths: List[str],
                 tgt_units: List[List[int]], src_n_frames: List[int], tgt_n_frames: List[int], ids: Optional[List[str]] = None,
                 tgt_dict=None, shuffle: bool = False, cfg=None):
        self.split, self.cfg, self.tgt_dict = split, cfg, tgt_dict
        self.src_n_frames, self.tgt_n_frames = list(src_n_frames), list(tgt_n_frames)
        self.n_samples = len(audio_paths)
        self.audio_paths, self.tgt_feat_paths, self.tgt_units, self.ids = audio_paths, tgt_feat_paths, tgt_units, ids
        assert self.n_samples == len(self.src_n_frames) == len(self.tgt_n_frames)
        assert ids is None or len(ids) == self.n_samples
        self.shuffle = bool(shuffle) if is_train_split else False   # :75
        self._lut = None

    # -- the run-length reduction (:92-113); vectorised, same three outputs (index_to_keep as a LongTensor)
    @staticmethod
    def _reduce_tgt(tokens):
        import torch
        t = np.asarray(tokens, dtype=np.int64)
        if t.size == 0:
            return [], [1], torch.zeros(0, dtype=torch.long)   # the unconditional append of :112
        start = np.ones(t.size, dtype=bool)
        start[1:] = t[1:] != t[:-1]
        keep = np.flatnonzero(start)
        dur = np.diff(np.append(keep, t.size))
        return t[keep].tolist(), dur.tolist(), torch.from_numpy(keep).long()

    def _encode(self, units) -> "torch.Tensor":
        """Dictionary.encode_line(" ".join(units), add_if_not_exist=False, append_eos=False).long() (:130-140): symbol
        str(k) sits at index k + nspecial; anything outside the dictionary maps to <unk>."""
        import torch
        d = self.tgt_dict
        u = np.asarray(units, dtype=np.int64)
        if self._lut is None:
            n = len(d) - d.nspecial if hasattr(d, "nspecial") else len(d) - 4
            self._lut = np.array([d.index(str(k)) for k in range(n)], dtype=np.int64)
        out = np.full(u.shape, d.unk(), dtype=np.int64)
        ok = (u >= 0) & (u < len(self._lut))
        out[ok] = self._lut[u[ok]]
        return torch.from_numpy(out)

    def __getitem__(self, index: int) -> ReprToReprDatasetItem:
        import torch
        src_feat = torch.from_numpy(np.load(self.audio_paths[index])).float()
        tgt_feat = torch.from_numpy(np.load(self.tgt_feat_paths[index])).float()
        units = self.tgt_units[index]
        reduced, _dur, keep = self._reduce_tgt(units)
        return ReprToReprDatasetItem(index=index, src_feat=src_feat, tgt_feat=tgt_feat, tgt_unit=self._encode(units),
                                     reduce_tgt_unit=self._encode(reduced), reduce_tgt_feat=tgt_feat[keep])

    def __len__(self):
        return self.n_samples

    def num_tokens(self, index):
        return self.tgt_n_frames[index]

    def size(self, index):
        return self.tgt_n_frames[index]

    @property
    def sizes(self):
        return np.array(self.tgt_n_frames)

    @property
    def can_reuse_epoch_itr_across_epochs(self):
        return True

    def ordered_indices(self):
        """:177-184: descending target length, ties in original (or, for shuffled training, random) order."""
        order = [np.random.permutation(len(self))] if self.shuffle else [np.arange(len(self))]
        order.append([-n for n in self.tgt_n_frames])
        return np.lexsort(order)

    def batch_sampler(self, max_tokens: int = 0, max_sentences: int = 0, bsz_mult: int = 1) -> List[np.ndarray]:
        """Batches of `ordered_indices()` under a padded-token budget (what fairseq's EpochBatchIterator asks
        `batch_by_size` for), through the native batcher."""
        idx = self.ordered_indices()
        sizes = self.sizes[idx]
        return [idx[s:e] for s, e in batch_by_size(sizes, max_tokens, max_sentences, bsz_mult)]

    def collater(self, samples: List[ReprToReprDatasetItem], return_order: bool = False) -> Dict:
        """:193-258."""
        import torch
        if len(samples) == 0:
            return {}
        B = len(samples)
        indices = torch.tensor([x.index for x in samples], dtype=torch.long)
        src_len = torch.tensor([x.src_feat.shape[0] for x in samples], dtype=torch.long)
        tgt_len = torch.tensor([x.tgt_feat.shape[0] for x in samples], dtype=torch.long)
        red_len = torch.tensor([x.reduce_tgt_unit.shape[0] for x in samples], dtype=torch.long)
        C = samples[0].src_feat.shape[1]
        src = samples[0].src_feat.new_zeros(B, int(src_len.max()), C)
        tgt = samples[0].src_feat.new_zeros(B, int(tgt_len.max()), C)
        tgt_unit = samples[0].tgt_unit.new_zeros(B, int(tgt_len.max()))
        red_unit = samples[0].reduce_tgt_unit.new_zeros(B, int(red_len.max()))
        red_feat = samples[0].reduce_tgt_feat.new_zeros(B, int(red_len.max()), C)
        for i, x in enumerate(samples):
            src[i, : x.src_feat.shape[0]] = x.src_feat
            tgt[i, : x.tgt_feat.shape[0]] = x.tgt_feat
            tgt_unit[i, : x.tgt_unit.shape[0]] = x.tgt_unit
            red_unit[i, : x.reduce_tgt_unit.shape[0]] = x.reduce_tgt_unit
            red_feat[i, : x.reduce_tgt_feat.shape[0]] = x.reduce_tgt_feat
        src_len, order = src_len.sort(descending=True)   # rows re-ordered by SOURCE length (:228)
        pick = lambda t: t.index_select(0, order)
        red_len = pick(red_len)
        return {
            "id": pick(indices),
            "net_input": {"src_tokens": pick(src), "src_lengths": src_len, "prev_output_tokens": None, "tgt_speaker": None},
            "speaker": None,
            "target": pick(tgt), "target_unit": pick(tgt_unit),
            "reduce_target": pick(red_feat), "reduce_target_unit": pick(red_unit),
            "target_lengths": pick(tgt_len), "reduce_target_lengths": red_len,
            "ntokens": red_len.sum().item(), "nsentences": B,
        }

    @classmethod
    def from_samples(cls, split: str, is_train_split: bool, samples: List[Dict], tgt_dict, shuffle: bool = False, cfg=None):
        """ReprToReprUnitDatasetCreator._from_list (:274-304)."""
        return cls(split, is_train_split, [s["src_audio"] for s in samples], [s["tgt_audio"] for s in samples],
                   [s["tgt_unit"] for s in samples], [int(s["src_n_frames"]) for s in samples],
                   [int(s["tgt_n_frames"]) for s in samples], ids=[s["id"] for s in samples], tgt_dict=tgt_dict,
                   shuffle=shuffle, cfg=cfg)

    @classmethod
    def from_tsv(cls, src_feat_dir: str, tgt_feat_dir: str, audio_root: str, splits: str, is_train_split: bool, tgt_dict,
                 shuffle: bool = False, cfg=None, **_unused):
        """ReprToReprUnitDatasetCreator.from_tsv (:371-399); several comma-separated splits concatenate."""
        parts = [cls.from_samples(sp, is_train_split, load_samples_from_tsv(src_feat_dir, tgt_feat_dir, audio_root, sp),
                                  tgt_dict, shuffle, cfg) for sp in splits.split(",")]
        if len(parts) == 1:
            return parts[0]
        samples_args = [sum((getattr(p, a) for p in parts), []) for a in
                        ("audio_paths", "tgt_feat_paths", "tgt_units", "src_n_frames", "tgt_n_frames", "ids")]
        return cls(splits, is_train_split, *samples_args[:5], ids=samples_args[5], tgt_dict=tgt_dict, shuffle=shuffle, cfg=cfg)

    @classmethod
    def from_manifest(cls, args, split: str, tgt_dict):
        """What `task.load_dataset(split)` calls (speech_decoder_task.py:159-170): directories from the task's flags,
        `shuffle` from the `--dummy-config` YAML (S2SDataConfig.shuffle, data_cfg.py)."""
        shuffle = False
        cfg_path = getattr(args, "dummy_config", None)
        if cfg_path and os.path.isfile(cfg_path):
            import yaml
            with open(cfg_path) as f:
                shuffle = bool((yaml.safe_load(f) or {}).get("shuffle", False))
        return cls.from_tsv(args.src_feat_dir, args.tgt_feat_dir, args.data, split, split.startswith("train"), tgt_dict,
                            shuffle=shuffle)


def write_tsv(path: str, lines: Iterable[str]):
    with open(path, "w") as f:
        f.write(HEADER + "\n")
        for ln in lines:
            f.write(ln + "\n")
