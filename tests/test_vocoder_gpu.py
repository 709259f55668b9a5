"""The unit vocoder (the step after the pass) on the GPU against the waveforms the UNTOUCHED reference modules
produced (tests/golden/vocoder_code_hifigan.npz, minted by oracle/make_golden.py --vocoder-only from codehifigan.py /
hifigan.py / fastspeech2.VariancePredictor) and against the pinned oracle on longer inputs.

Tolerances: durations are integers: bit-exact.  Waveform: fp32 kernels against an fp32 reference through ~60 convolutions
of up to 5632 products each: max-abs <= 2e-4 on a signal in [-1, 1] (measured ~2e-5)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from diffnorm_b200.vocoder import CodeHiFiGANVocoder, load_code, process_units  # noqa: E402
from oracle import vocoder_oracle as V  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_vocoder_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "vocoder_code_hifigan.npz"))
    sd = V.init_state_dict(int(g["weight_seed"]))
    voc = CodeHiFiGANVocoder({"generator": sd})
    for name, dp in (("dur", True), ("nodur", False), ("reduced", True)):
        code = torch.from_numpy(g[f"{name}_code"]).view(1, -1)
        wav = voc({"code": code.cuda()}, dur_prediction=dp)
        assert voc.last_durations.cpu().tolist() == g[f"{name}_dur"].tolist()           # integers: bit-exact
        want = torch.from_numpy(g[f"{name}_wave"])
        assert wav.shape == want.shape and wav.numel() == 320 * int(g[f"{name}_dur"].sum())
        d = float((wav.cpu() - want).abs().max())
        print(f"[parity] vocoder {name}: {wav.numel()} samples, max|d| {d:.2e} (std {float(want.std()):.3f})")
        assert d < 2e-4


def test_vocoder_long_utterance_against_oracle_and_weight_norm_free_checkpoint(tmp_path):
    sd = V.init_state_dict(3)
    rng = np.random.default_rng(4)
    units = rng.integers(0, 1000, size=700)
    units[[5, 80]] = -1                                     # invalid codes are dropped by the wrapper (vocoder.py:234-235)
    code = torch.from_numpy(units)
    want, dur = V.code_to_waveform(sd, code, dur_prediction=True)
    # a checkpoint with weight norm already removed (plain .weight keys) must load and give the same result
    plain = {}
    for k, v in sd.items():
        if k.endswith(".weight_v"):
            plain[k[:-2]] = V._w(sd, k[:-len(".weight_v")])
        elif not k.endswith(".weight_g"):
            plain[k] = v
    torch.save({"generator": plain}, tmp_path / "g.pt")
    for ckpt in ({"generator": sd}, str(tmp_path / "g.pt")):
        voc = CodeHiFiGANVocoder(ckpt)
        wav = voc({"code": code.view(1, -1)}, dur_prediction=True)
        assert voc.last_durations.cpu().tolist() == dur.tolist()
        d = float((wav.cpu() - want).abs().max())
        print(f"[parity] vocoder 698 units -> {wav.numel()} samples: max|d| {d:.2e}")
        assert wav.numel() == 320 * int(dur.sum()) and d < 2e-4
    assert voc({"code": torch.full((1, 4), -1)}).numel() == 0


def test_code_file_reader(tmp_path):
    (tmp_path / "c.txt").write_text("a|3 3 4 4 4 3\nb|7\n")
    assert load_code(str(tmp_path / "c.txt"), reduce=True) == [[3, 4, 3], [7]]
    assert load_code(str(tmp_path / "c.txt"), reduce=False) == [[3, 3, 4, 4, 4, 3], [7]]
    assert process_units([1, 1, 2], reduce=True) == [1, 2]
