#!/usr/bin/env python
"""Throughput of the k-means unit quantiser on resident features vs scikit-learn on the host cores."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from diffnorm_b200.kmeans import KMeansQuantizer  # noqa: E402

rng = np.random.default_rng(0)
K, D, N = 1000, 768, 64000
centers = rng.standard_normal((K, D)).astype(np.float32)
feats = rng.standard_normal((N, D)).astype(np.float32)
q = KMeansQuantizer(centers)
x = torch.from_numpy(feats).cuda()
for _ in range(3):
    q.predict(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    u = q.predict(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
from sklearn.cluster import KMeans  # noqa: E402
km = KMeans(n_clusters=K, n_init=1, max_iter=1).fit(feats[:2000])
km.cluster_centers_ = centers
t0 = time.perf_counter()
ref = km.predict(feats[:16000])
cpu_s = time.perf_counter() - t0
print(json.dumps({"config": "kmeans_quantize", "frames": N, "ms": ms, "frames_per_s": N / (ms * 1e-3),
                  "sklearn_frames_per_s": 16000 / cpu_s, "cores": os.cpu_count(),
                  "agreement_with_sklearn_16k": float((u.cpu().numpy()[:16000] == ref).mean())}))
