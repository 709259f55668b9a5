"""Command-line unit quantiser: drop-in for the dumped-features mode of
examples/textless_nlp/gslm/speech2unit/clustering/quantize_with_kmeans.py (:80-121) — same flags, same
``{basename}|{units}`` output — with the nearest-centroid search on the GPU (diffnorm_b200/kmeans.py).

    python -m diffnorm_b200.quantize_cli --kmeans_model_path km.bin --manifest_path split.manifest.tsv \
        --out_quantized_file_path out/split.quant.tsv

The manifest is the feature manifest of SURVEY Appendix B (line 1 = feature dir, then ``{id}.feat.npy \\t N``); extracting
mHuBERT features from audio stays with the reference (out of scope)."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from .data import read_manifest
from .kmeans import KMeansQuantizer, load_centers, write_quantized


def cli_main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--kmeans_model_path", type=str, required=True, help="joblib-pickled scikit-learn k-means model, or a .npy of centroids")
    p.add_argument("--manifest_path", type=str, default=None, help="feature manifest (root dir + '{id}.feat.npy\\tN' rows)")
    p.add_argument("--features_path", type=str, default=None, help="one .npy of features [N, 768] (single pseudo-utterance)")
    p.add_argument("--out_quantized_file_path", type=str, required=True)
    a = p.parse_args(argv)
    centers = np.load(a.kmeans_model_path) if a.kmeans_model_path.endswith(".npy") else load_centers(a.kmeans_model_path)
    q = KMeansQuantizer(centers)
    if a.features_path is not None:
        names, feats = [os.path.basename(a.features_path)], [np.load(a.features_path)]
    else:
        root, rows = read_manifest(a.manifest_path)
        names = [r[0] for r in rows]
        feats = [np.load(os.path.join(root, n)) for n in names]
    units = q.predict_many(feats)
    write_quantized(a.out_quantized_file_path, names, units)
    print(f"Wrote {len(names)} quantized utterances to {a.out_quantized_file_path}")


if __name__ == "__main__":
    cli_main(sys.argv[1:])
