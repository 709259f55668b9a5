"""K-means unit quantisation: 768-d mHuBERT features -> index of the nearest of 1000 centroids, the
step that produces the "original units" fed to `_reduce_tgt` (examples/textless_nlp/gslm/speech2unit/clustering/
quantize_with_kmeans.py:109-121: ``kmeans_model.predict(feats)`` of a joblib-pickled scikit-learn model).

argmin_k ||x - c_k||^2 = argmax_k (x . c_k - 0.5 ||c_k||^2): one tensor-core GEMM + the warp-shuffle argmax already on the
path.  The contraction must resolve near-ties between centroids, so both operands are split into bf16 hi + lo parts and the
GEMM runs over K = 3 * 768 ([hi|hi|lo] x [hi|lo|hi]); the -0.5 ||c||^2 term is an exact fp32 bias.  First minimum wins, like
numpy / scikit-learn's argmin.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import _lib, ops
from .ops import GemmPlan
from .packing import BK, WT, rup

bf16, f32 = torch.bfloat16, torch.float32


class KMeansQuantizer:
    def __init__(self, centers, device: str = "cuda"):
        """centers: [K, D] float array / tensor (``kmeans_model.cluster_centers_``)."""
        if not torch.cuda.is_available():
            raise RuntimeError("KMeansQuantizer needs a CUDA device: the product has no CPU path")
        c = torch.as_tensor(np.asarray(centers), dtype=torch.float64)
        self.K, self.D = c.shape
        if self.D % 64:
            raise ValueError("feature dimension must be a multiple of 64")
        self.dev = torch.device(device)
        c32 = c.float()
        hi = c32.to(bf16).float()
        lo = (c32 - hi).to(bf16).float()
        Kp = rup(self.K, 16)
        W = torch.zeros(Kp, 3 * self.D)
        W[: self.K] = torch.cat([hi, lo, hi], dim=1)
        bias = torch.full((Kp,), -3.0e38)                       # pad classes can never win
        bias[: self.K] = (-0.5 * (c * c).sum(dim=1)).float()    # exact: computed in float64 from the given centers
        self.plan = GemmPlan(W.to(bf16).contiguous().to(self.dev), [(0, 0, 3 * self.D // BK, 0, 0)], Kp, (Kp + WT - 1) // WT,
                             _lib.EPI_F32, bias=bias.to(self.dev), name="kmeans.scores")
        self.Kp = Kp

    @torch.no_grad()
    def predict(self, feats: torch.Tensor) -> torch.Tensor:
        """feats fp32 [N, D] on the device -> int64 [N] centroid indices."""
        x = feats.to(self.dev, dtype=f32).contiguous()
        n = x.shape[0]
        a = ops.split_bf16x3(x)
        scores = torch.empty(n, self.Kp, dtype=f32, device=self.dev)
        self.plan.run(a, scores, 1, n)
        return ops.argmax_units(scores, self.K, 0)

    def predict_many(self, feats: Sequence[np.ndarray], max_rows: int = 1 << 18) -> List[np.ndarray]:
        """Utterance-batched form of the reference loop (:113-121): packs utterances up to `max_rows` frames per launch."""
        out: List[np.ndarray] = [None] * len(feats)
        i = 0
        while i < len(feats):
            j, rows = i, 0
            while j < len(feats) and (rows == 0 or rows + len(feats[j]) <= max_rows):
                rows += len(feats[j])
                j += 1
            packed = torch.from_numpy(np.concatenate([np.asarray(f, dtype=np.float32) for f in feats[i:j]], axis=0)).pin_memory()
            units = self.predict(packed.to(self.dev, non_blocking=True)).cpu().numpy()
            o = 0
            for k in range(i, j):
                out[k] = units[o:o + len(feats[k])]
                o += len(feats[k])
            i = j
        return out


def load_centers(kmeans_model_path: str) -> np.ndarray:
    """``joblib.load`` of the reference's pickled scikit-learn model -> cluster_centers_ (:104-106)."""
    import joblib
    with open(kmeans_model_path, "rb") as f:
        model = joblib.load(f)
    return np.asarray(model.cluster_centers_)


def write_quantized(path: str, names: Sequence[str], units: Sequence[np.ndarray]):
    """Output format of quantize_with_kmeans.py:116-121: ``{basename}|{space separated units}``."""
    import os
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        for n, u in zip(names, units):
            f.write(f"{os.path.basename(n)}|{' '.join(str(int(p)) for p in u)}\n")
