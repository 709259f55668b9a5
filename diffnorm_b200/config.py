"""Shape description of the DiffNorm modules on the path (LM:709-807 denoiser, LM:1035-1097 VAE)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

VOCAB = 1004       # LM:1095
UNIT_OFFSET = 4    # LM:1451  (fairseq Dictionary specials <s>,<pad>,</s>,<unk>)
TIMESTEPS = 200    # diff_discrete.py:84


@dataclass
class DiffNormConfig:
    latent_dim: int = 16
    feat_dim: int = 768
    hid: int = 512
    depth: int = 12
    heads: int = 8
    dim_head: int = 64
    wn_stacks: int = 4
    wn_layers: int = 8
    cond_mult: int = 4
    vae_depth: int = 6
    vae_heads: int = 8
    vae_dim_head: int = 96
    vae_stacks: int = 2
    vae_layers: int = 3
    vocab: int = VOCAB
    timesteps: int = TIMESTEPS
    chan_mults: List[int] = field(default_factory=list)

    def __post_init__(self):
        if not self.chan_mults:  # LM:1044-1051
            self.chan_mults = {16: [4, 3, 2], 32: [4, 3], 128: [3]}[self.latent_dim]

    @property
    def dim_time(self) -> int:
        return self.hid * self.cond_mult

    @staticmethod
    def ff_inner(dim: int) -> int:  # LM:888
        return int(dim * 4 * 2 / 3)

    def enc_widths(self) -> List[Tuple[int, int]]:
        out, cur = [], self.feat_dim
        for m in self.chan_mults:
            out.append((cur, cur // m))
            cur //= m
        return out

    def dec_widths(self) -> List[Tuple[int, int]]:
        out, cur, first = [], self.enc_widths()[-1][1], True
        for m in reversed(self.chan_mults):
            tgt = cur * m
            if first:
                cur, first = cur // 2, False
            out.append((cur, tgt))
            cur = tgt
        return out

    @classmethod
    def from_state_dict(cls, sd: Dict[str, "object"]) -> "DiffNormConfig":
        """Infer the shapes from a reference LatentDiscreteModel state_dict (keys model.* / speech_decoder.*)."""
        hid, z = sd["model.init_conv.weight"].shape[:2]
        depth = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("model.transformer.layers."))
        wn_stacks = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("model.wavenet.stacks."))
        wn_layers = 1 + max(int(k.split(".")[5]) for k in sd if k.startswith("model.wavenet.stacks.0.blocks."))
        dim_time = sd["model.to_time_cond.1.weight"].shape[0]
        vdepth = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("speech_decoder.decoder_tf.layers."))
        feat = sd["speech_decoder.decoder_lm.weight"].shape[1]
        vocab = sd["speech_decoder.decoder_lm.weight"].shape[0]
        return cls(latent_dim=int(z), feat_dim=int(feat), hid=int(hid), depth=depth, wn_stacks=wn_stacks,
                   wn_layers=wn_layers, cond_mult=dim_time // int(hid), vae_depth=vdepth, vocab=int(vocab))
