"""CPU restatement of the step AFTER the normalization pass: the duration-aware unit vocoder.

TEST INFRASTRUCTURE ONLY — nothing under diffnorm_b200/ imports this file, and no product code for this stage exists
yet: this oracle (pinned to the live reference through tests/golden/vocoder_*.npz, minted by oracle/make_golden.py
--vocoder-only) is the first half of the next widening step; the CUDA path is the next round's.

What it restates, from the reference files it follows:
  * ``CodeGenerator.forward``                 fairseq/models/text_to_speech/codehifigan.py:49-76
      code [1, T] -> embedding [1, 128, T] -> (dur_prediction) VariancePredictor -> dur = clamp(round(exp(log_dur) - 1), 1)
      -> repeat_interleave along T -> HiFi-GAN generator -> waveform [1, 1, 320 * sum(dur)]
  * ``VariancePredictor.forward``             fairseq/models/text_to_speech/fastspeech2.py:117-151  (eval: dropout off)
      conv1d(k 3, pad 1) + ReLU -> LayerNorm -> conv1d(k 3, pad 1) + ReLU -> LayerNorm -> Linear(128, 1)
  * ``Generator.forward`` / ``ResBlock``      fairseq/models/text_to_speech/hifigan.py:20-179
      conv_pre(k 7) -> 5 x [leaky_relu 0.1 -> ConvTranspose1d(k, stride u, pad (k-u)/2) -> mean of 3 ResBlocks(k in 3,7,11;
      dilations 1,3,5)] -> leaky_relu(0.01) -> conv_post(k 7) -> tanh.  Weights carry weight-norm (w = g v / ||v|| over all but
      dim 0); ``CodeHiFiGANVocoder`` strips it after loading (vocoder.py:215-229), which changes no value.
  * the driver's unit handling                examples/speech_to_speech/generate_waveform_from_code.py:33-52, vocoder.py:231-237
      ``process_units(reduce=True)`` = consecutive-duplicate removal; codes < 0 are dropped before the model.

The configuration is the published unit-vocoder one (mHuBERT 1000 units, 16 kHz, hop 320 = 5*4*4*2*2).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

VOCODER_CFG = {
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
LRELU_SLOPE = 0.1            # hifigan.py:7
HOP = 320


def process_units(units: List[int], reduce: bool = False) -> List[int]:
    """generate_waveform_from_code.py:33-38."""
    if not reduce:
        return list(units)
    return [u for i, u in enumerate(units) if i == 0 or u != units[i - 1]]


def weight_norm_keys(cfg=VOCODER_CFG) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every tensor of CodeGenerator.state_dict(), in module order (hifigan.py:113-150, codehifigan.py:9-25)."""
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def wn(prefix, cout, cin, k, transposed=False):
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        out.extend([(prefix + ".bias", (cout,)), (prefix + ".weight_g", (shape[0], 1, 1)), (prefix + ".weight_v", shape)])

    c0 = cfg["upsample_initial_channel"]
    wn("conv_pre", c0, cfg["model_in_dim"], 7)
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        wn(f"ups.{i}", c0 // 2 ** (i + 1), c0 // 2 ** i, k, transposed=True)
    nk = len(cfg["resblock_kernel_sizes"])
    for i in range(len(cfg["upsample_rates"])):
        ch = c0 // 2 ** (i + 1)
        for j, k in enumerate(cfg["resblock_kernel_sizes"]):
            for grp in ("convs1", "convs2"):
                for d in range(3):
                    wn(f"resblocks.{i * nk + j}.{grp}.{d}", ch, ch, k)
    wn("conv_post", 1, c0 // 2 ** len(cfg["upsample_rates"]), 7)
    out.append(("dict.weight", (cfg["num_embeddings"], cfg["embedding_dim"])))
    p = cfg["dur_predictor_params"]
    h, e, k = p["var_pred_hidden_dim"], p["encoder_embed_dim"], p["var_pred_kernel_size"]
    out += [("dur_predictor.conv1.0.weight", (h, e, k)), ("dur_predictor.conv1.0.bias", (h,)),
            ("dur_predictor.ln1.weight", (h,)), ("dur_predictor.ln1.bias", (h,)),
            ("dur_predictor.conv2.0.weight", (h, h, k)), ("dur_predictor.conv2.0.bias", (h,)),
            ("dur_predictor.ln2.weight", (h,)), ("dur_predictor.ln2.bias", (h,)),
            ("dur_predictor.proj.weight", (1, h)), ("dur_predictor.proj.bias", (1,))]
    return out


def init_state_dict(seed: int, cfg=VOCODER_CFG) -> Dict[str, torch.Tensor]:
    """Seeded weights with the reference's key layout.  NOT the reference's init law (its N(0, 0.01) convs give a near-silent
    waveform and exp(log_dur) - 1 ~ 0 everywhere, so every duration would clamp to 1): here the weight-norm gains keep the
    signal at O(1) through the five stages (row norm of the effective weight = g; a transposed conv of stride u needs
    g ~ sqrt(u / 2)), the output sits in tanh's non-saturated range, and the duration head spreads over ~1..6 frames, so the
    parity checks are not vacuous."""
    gen = torch.Generator().manual_seed(seed)
    rn = lambda *shape: torch.randn(*shape, generator=gen)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in weight_norm_keys(cfg):
        if name.endswith(".weight_v"):
            sd[name] = rn(*shape)
        elif name.endswith(".weight_g"):
            if name.startswith("ups."):
                u = cfg["upsample_rates"][int(name.split(".")[1])]
                g0 = 1.3 * (u / 2.0) ** 0.5
            elif name.startswith("conv_post"):
                g0 = 0.35
            elif ".convs2." in name:
                g0 = 0.5
            else:
                g0 = 1.3
            sd[name] = g0 * (1.0 + 0.1 * rn(*shape))
        elif name in ("dur_predictor.ln1.weight", "dur_predictor.ln2.weight"):
            sd[name] = 1.0 + 0.1 * rn(*shape)
        elif name == "dict.weight":
            sd[name] = rn(*shape)
        elif name == "dur_predictor.proj.weight":
            sd[name] = rn(*shape) * (0.6 / shape[1] ** 0.5)
        elif name == "dur_predictor.proj.bias":
            sd[name] = torch.full(shape, 1.1)
        elif name.endswith(".weight"):                      # duration-predictor convs
            sd[name] = rn(*shape) * (2.0 / (shape[1] * shape[2])) ** 0.5
        else:
            sd[name] = 0.05 * rn(*shape)
    return sd


def _w(sd, prefix: str) -> torch.Tensor:
    """torch.nn.utils.weight_norm (dim 0): w = g * v / ||v||_{all dims but 0} (hifigan.py:28-74 wraps every conv)."""
    v, g = sd[prefix + ".weight_v"], sd[prefix + ".weight_g"]
    return g * v / v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))


def get_padding(k: int, d: int = 1) -> int:   # hifigan.py:16-17
    return (k * d - d) // 2


def duration_predictor(sd, x: torch.Tensor) -> torch.Tensor:
    """fastspeech2.py:145-151 in eval mode.  x [B, T, C] -> log-durations [B, T]."""
    p = "dur_predictor."
    h = F.relu(F.conv1d(x.transpose(1, 2), sd[p + "conv1.0.weight"], sd[p + "conv1.0.bias"], padding=1)).transpose(1, 2)
    h = F.layer_norm(h, h.shape[-1:], sd[p + "ln1.weight"], sd[p + "ln1.bias"])
    h = F.relu(F.conv1d(h.transpose(1, 2), sd[p + "conv2.0.weight"], sd[p + "conv2.0.bias"], padding=1)).transpose(1, 2)
    h = F.layer_norm(h, h.shape[-1:], sd[p + "ln2.weight"], sd[p + "ln2.bias"])
    return F.linear(h, sd[p + "proj.weight"], sd[p + "proj.bias"]).squeeze(2)


def resblock(sd, prefix: str, x: torch.Tensor, k: int, dilations) -> torch.Tensor:
    """hifigan.py:91-98."""
    for i, d in enumerate(dilations):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, _w(sd, f"{prefix}.convs1.{i}"), sd[f"{prefix}.convs1.{i}.bias"], dilation=d, padding=get_padding(k, d))
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, _w(sd, f"{prefix}.convs2.{i}"), sd[f"{prefix}.convs2.{i}.bias"], padding=get_padding(k, 1))
        x = xt + x
    return x


def generator(sd, x: torch.Tensor, cfg=VOCODER_CFG) -> torch.Tensor:
    """hifigan.py:152-168.  x [B, 128, T'] -> waveform [B, 1, 320 T']."""
    nk = len(cfg["resblock_kernel_sizes"])
    x = F.conv1d(x, _w(sd, "conv_pre"), sd["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, _w(sd, f"ups.{i}"), sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, (rk, rd) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            r = resblock(sd, f"resblocks.{i * nk + j}", x, rk, rd)
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x)                       # default slope 0.01 (hifigan.py:164)
    x = F.conv1d(x, _w(sd, "conv_post"), sd["conv_post.bias"], padding=3)
    return torch.tanh(x)


@torch.no_grad()
def code_to_waveform(sd, code: torch.Tensor, dur_prediction: bool = True, cfg=VOCODER_CFG):
    """vocoder.py:231-237 + codehifigan.py:49-76: code int64 [T] (entries < 0 dropped) -> (waveform [320 sum(dur)], dur [T'])."""
    code = code[code >= 0].view(1, -1)
    x = F.embedding(code, sd["dict.weight"]).transpose(1, 2)              # [1, 128, T]
    dur = torch.ones(code.shape[1], dtype=torch.long)
    if dur_prediction:
        log_dur = duration_predictor(sd, x.transpose(1, 2))
        dur = torch.clamp(torch.round(torch.exp(log_dur) - 1).long(), min=1).view(-1)
        x = torch.repeat_interleave(x, dur, dim=2)
    return generator(sd, x, cfg).squeeze(), dur
