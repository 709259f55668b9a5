#!/usr/bin/env python
"""Unit vocoder throughput: reduced units -> waveform, one utterance at a time like the reference driver.
Reports audio seconds per second and the fp32 algorithmic FLOP rate (155.6 M MAC per duration frame of 20 ms)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from diffnorm_b200.vocoder import VOCODER_CFG, CodeHiFiGANVocoder  # noqa: E402
from oracle import vocoder_oracle as V  # noqa: E402


def mac_per_frame(c=VOCODER_CFG):
    ch, up, tot = c["upsample_initial_channel"], 1, 7 * c["model_in_dim"] * c["upsample_initial_channel"]
    for u, k in zip(c["upsample_rates"], c["upsample_kernel_sizes"]):
        tot += up * ch * (ch // 2) * k          # transposed conv: k/u taps per output, u outputs per input
        ch //= 2
        up *= u
        tot += up * sum(rk * 2 * 3 for rk in c["resblock_kernel_sizes"]) * ch * ch
    return tot + up * 7 * ch


def main():
    sd = V.init_state_dict(3)
    voc = CodeHiFiGANVocoder({"generator": sd})
    rng = np.random.default_rng(0)
    n = int(os.environ.get("UNITS", 600))
    code = torch.from_numpy(rng.integers(0, 1000, size=n)).view(1, -1).cuda()
    for dp in (True, False):
        wav = voc({"code": code}, dur_prediction=dp)
        torch.cuda.synchronize()
        frames = wav.numel() // 320
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            voc({"code": code}, dur_prediction=dp)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        t0 = time.perf_counter()
        with torch.no_grad():
            V.code_to_waveform(sd, code.cpu().view(-1)[:100], dur_prediction=dp)
        cpu_s = time.perf_counter() - t0
        print(json.dumps({"units": n, "dur_prediction": dp, "frames": frames, "audio_s": frames * 0.02, "ms": ms,
                          "audio_s_per_s": frames * 0.02 / (ms * 1e-3), "fp32_tflops": 2 * mac_per_frame() * frames / (ms * 1e-3) / 1e12,
                          "oracle_cpu_s_for_100_units": cpu_s}))


if __name__ == "__main__":
    main()
