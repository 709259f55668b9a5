"""The step after the normalization pass: the duration-aware unit vocoder that turns the reduced,
normalized units back into a waveform — drop-in counterpart of the reference's ``CodeHiFiGANVocoder``
(fairseq/models/text_to_speech/vocoder.py:213-245) over ``CodeGenerator`` (codehifigan.py:9-76, hifigan.py:99-179) and the
``VariancePredictor`` duration head (fastspeech2.py:117-151), as driven by
examples/speech_to_speech/generate_waveform_from_code.py:78-96.

Same constructor arguments, ``forward(x, dur_prediction)`` signature, state_dict keys (``generator`` checkpoints with or
without weight norm load as they are) and output (a 1-D fp32 waveform of 320 samples per duration frame); all arithmetic runs
in the fp32 CUDA kernels of csrc/vocoder.cu through the C ABI (dn_voc_*).  No torch / CPU implementation exists behind it.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from ._lib import check, lib

VOCODER_CFG = {   # the published mHuBERT-1000 unit vocoder (16 kHz, hop 320 = 5*4*4*2*2)
    "resblock_kernel_sizes": [3, 7, 11],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    "upsample_rates": [5, 4, 4, 2, 2],
    "upsample_kernel_sizes": [11, 8, 8, 4, 4],
    "upsample_initial_channel": 512,
    "model_in_dim": 128,
    "num_embeddings": 1000,
    "embedding_dim": 128,
    "dur_predictor_params": {"encoder_embed_dim": 128, "var_pred_hidden_dim": 128, "var_pred_kernel_size": 3,
                             "var_pred_dropout": 0.5},
}
LRELU_SLOPE = 0.1   # hifigan.py:7
f32, i64 = torch.float32, torch.int64


def _p(t):
    return None if t is None else t.data_ptr()


def _st():
    return torch.cuda.current_stream().cuda_stream


def process_units(units, reduce: bool = False):
    """generate_waveform_from_code.py:33-38: optional removal of consecutive duplicates."""
    if not reduce:
        return units
    return [u for i, u in enumerate(units) if i == 0 or u != units[i - 1]]


def load_code(in_file: str, reduce: bool, filter_score: Optional[float] = None):
    """generate_waveform_from_code.py:41-55: lines ``{sample_id}|{units}`` (what quantize_cli / the S2UT generator write)."""
    out = []
    with open(in_file) as f:
        for line in f:
            sample_id, units = line.strip().split("|")
            if filter_score is not None and float(sample_id.split("=")[1]) < filter_score:
                continue
            out.append(list(map(int, process_units(units.split(), reduce))))
    return out


class CodeHiFiGANVocoder:
    """vocoder.py:213-245.  ``checkpoint`` is a path to a ``{"generator": state_dict}`` file or that state_dict itself."""

    def __init__(self, checkpoint, model_cfg: Optional[Dict] = None, fp16: bool = False, device: str = "cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("CodeHiFiGANVocoder needs a CUDA device: the product has no CPU path")
        if fp16:
            raise NotImplementedError("the unit vocoder runs in fp32 (the reference's default)")
        self.cfg = dict(VOCODER_CFG if model_cfg is None else model_cfg)
        if self.cfg.get("multispkr") or self.cfg.get("f0"):
            raise NotImplementedError("speaker / f0 conditioning (codehifigan.py:28-47) is not on the DiffNorm path")
        sd = torch.load(checkpoint, map_location="cpu") if isinstance(checkpoint, str) else checkpoint
        sd = sd.get("generator", sd)
        self.dev = torch.device(device)
        self.dur_prediction_available = any(k.startswith("dur_predictor.") for k in sd)
        self.w: Dict[str, torch.Tensor] = {}
        for k, v in sd.items():
            if k.endswith(".weight_g"):
                continue
            if k.endswith(".weight_v"):   # fold weight norm (dim 0): w = g v / ||v|| — what remove_weight_norm leaves (vocoder.py:226)
                g = sd[k[:-2] + "_g"].float()
                vv = v.float()
                v = g * vv / vv.flatten(1).norm(dim=1).view(-1, *([1] * (vv.dim() - 1)))
                k = k[:-2]
            self.w[k] = v.float().contiguous().to(self.dev)
        self.multispkr = False
        self.model = self   # the driver reads vocoder.model.multispkr (generate_waveform_from_code.py:70)

    # ---- kernels ---------------------------------------------------------------------------------------------
    def _conv(self, x, name, K, dil=1, in_slope=1.0, act=0, res=None, scale=1.0, out=None, accumulate=False):
        w, b = self.w[name + ".weight"], self.w.get(name + ".bias")
        cout, cin = w.shape[0], w.shape[1]
        L = x.shape[1]
        assert x.shape[0] == cin and w.shape[2] == K
        if out is None:
            out = torch.empty(cout, L, dtype=f32, device=self.dev)
        check(lib.dn_voc_conv1d(_p(x), L, cin, _p(w), _p(b), cout, K, dil, dil * (K - 1) // 2, in_slope, act, _p(res), scale,
                                int(accumulate), _p(out), _st()), "dn_voc_conv1d")
        return out

    def _convt(self, x, name, K, stride, in_slope):
        w, b = self.w[name + ".weight"], self.w.get(name + ".bias")   # [Cin, Cout, K]
        cin, cout = w.shape[0], w.shape[1]
        L = x.shape[1]
        out = torch.empty(cout, L * stride, dtype=f32, device=self.dev)
        check(lib.dn_voc_conv_transpose1d(_p(x), L, cin, _p(w), _p(b), cout, K, stride, (K - stride) // 2, in_slope, _p(out), _st()),
              "dn_voc_conv_transpose1d")
        return out

    def _ln(self, x, name):
        out = torch.empty_like(x)
        check(lib.dn_voc_layernorm(_p(x), x.shape[0], x.shape[1], _p(self.w[name + ".weight"]), _p(self.w[name + ".bias"]), _p(out),
                                   _st()), "dn_voc_layernorm")
        return out

    # ---- model ------------------------------------------------------------------------------------------------
    def durations(self, code: torch.Tensor, dur_prediction: bool):
        """codehifigan.py:57-70.  code int64 [T] on the device -> (dur [T], start [T+1]) int64 on the device."""
        T = code.shape[0]
        log_dur = None
        if dur_prediction:
            if not self.dur_prediction_available:
                raise ValueError("this checkpoint has no duration predictor")
            ones = torch.arange(T + 1, dtype=i64, device=self.dev)
            x = torch.empty(self.cfg["embedding_dim"], T, dtype=f32, device=self.dev)
            check(lib.dn_voc_embed_repeat(_p(code), T, _p(self.w["dict.weight"]), x.shape[0], _p(ones), T, _p(x), _st()),
                  "dn_voc_embed_repeat")
            k = self.cfg["dur_predictor_params"]["var_pred_kernel_size"]
            h = self._ln(self._conv(x, "dur_predictor.conv1.0", k, act=1), "dur_predictor.ln1")
            h = self._ln(self._conv(h, "dur_predictor.conv2.0", k, act=1), "dur_predictor.ln2")
            pw = self.w["dur_predictor.proj.weight"].view(1, -1, 1)
            self.w.setdefault("dur_predictor.proj1.weight", pw.contiguous())
            self.w.setdefault("dur_predictor.proj1.bias", self.w["dur_predictor.proj.bias"])
            log_dur = self._conv(h, "dur_predictor.proj1", 1).view(-1)
        dur = torch.empty(T, dtype=i64, device=self.dev)
        start = torch.empty(T + 1, dtype=i64, device=self.dev)
        check(lib.dn_voc_durations(_p(log_dur), T, _p(dur), _p(start), _st()), "dn_voc_durations")
        return dur, start

    def generator(self, x: torch.Tensor) -> torch.Tensor:
        """hifigan.py:152-168.  x fp32 [128, L] -> waveform fp32 [320 L]."""
        c = self.cfg
        nk = len(c["resblock_kernel_sizes"])
        x = self._conv(x, "conv_pre", 7)
        for i, (u, k) in enumerate(zip(c["upsample_rates"], c["upsample_kernel_sizes"])):
            x = self._convt(x, f"ups.{i}", k, u, LRELU_SLOPE)
            xs = torch.empty_like(x)
            tmp, cur = torch.empty_like(x), [torch.empty_like(x), torch.empty_like(x)]
            for j, (rk, rd) in enumerate(zip(c["resblock_kernel_sizes"], c["resblock_dilation_sizes"])):
                p = f"resblocks.{i * nk + j}"
                r = x                                               # hifigan.py:91-98: x <- c2(lrelu(c1(lrelu(x)))) + x, three times
                for t, d in enumerate(rd):
                    self._conv(r, f"{p}.convs1.{t}", rk, dil=d, in_slope=LRELU_SLOPE, out=tmp)
                    last = t == len(rd) - 1
                    if last:    # the mean over the resblocks (hifigan.py:160-163) is folded into each one's last convolution
                        self._conv(tmp, f"{p}.convs2.{t}", rk, in_slope=LRELU_SLOPE, res=r, scale=1.0 / nk, out=xs, accumulate=j > 0)
                    else:
                        r = self._conv(tmp, f"{p}.convs2.{t}", rk, in_slope=LRELU_SLOPE, res=r, out=cur[t & 1])
            x = xs
        return self._conv(x, "conv_post", 7, in_slope=0.01, act=2).view(-1)   # F.leaky_relu default slope, then tanh (:164-166)

    @torch.no_grad()
    def forward(self, x: Dict[str, torch.Tensor], dur_prediction: bool = False) -> torch.Tensor:
        """vocoder.py:231-237: x["code"] int64 [1, T]; entries < 0 are invalid and dropped.  Returns the waveform [320 sum(dur)]."""
        assert "code" in x
        code = x["code"].to(self.dev).view(-1)
        code = code[code >= 0].contiguous()
        T = int(code.shape[0])
        if T == 0:
            return torch.zeros(0, dtype=f32, device=self.dev)
        dur, start = self.durations(code, dur_prediction)
        Lo = int(start[-1].item())          # the one host read: the output length decides the allocation (as in the reference)
        emb = torch.empty(self.cfg["embedding_dim"], Lo, dtype=f32, device=self.dev)
        check(lib.dn_voc_embed_repeat(_p(code), T, _p(self.w["dict.weight"]), emb.shape[0], _p(start), Lo, _p(emb), _st()),
              "dn_voc_embed_repeat")
        self.last_durations = dur
        return self.generator(emb)

    __call__ = forward

    def cuda(self):
        return self
