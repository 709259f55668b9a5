"""`diff_discrete` model + arch (reference: fairseq/models/text_to_speech/diff_discrete.py:25-172)."""
from __future__ import annotations

import torch

from ..compat import FairseqEncoderModel, lengths_to_mask, register_model, register_model_architecture
from ..latent_module import LatentDiscreteModel
from .speech_vae_decoder import SpeechVAEDecoder, add_shared_args


@register_model("diff_discrete")
class DiffDiscreteModel(FairseqEncoderModel):
    def __init__(self, args, encoder):
        super().__init__(encoder)
        self.args = args

    def forward(self, target_feature, target_unit, **model_kwargs):
        src_mask = lengths_to_mask(model_kwargs["src_lengths"])
        tgt_mask = lengths_to_mask(model_kwargs["tgt_lengths"])
        return self.encoder(target_feature, target_unit, src_feature=model_kwargs["src_feature"], src_mask=src_mask,
                            tgt_mask=tgt_mask, unk_token=model_kwargs["unk_token"])  # loss dict, :42-55

    def get_normalized_probs(self, net_output, log_probs, sample=None):
        logits = net_output[0]
        return torch.log_softmax(logits, dim=-1) if log_probs else torch.softmax(logits, dim=-1)

    @classmethod
    def build_model(cls, args, task):
        """:70-85.  The frozen VAE comes from ``--speech_decoder_ckpt`` (a `speech_vae_decoder` checkpoint); when
        the flag is unset (synthetic / from-checkpoint use) a fresh VAE of the same shape is created and the caller's
        ``load_state_dict`` fills it (a diffusion checkpoint carries the VAE under ``encoder.speech_decoder.*``)."""
        ckpt = getattr(args, "speech_decoder_ckpt", None)
        vae = SpeechVAEDecoder.build_model(args, task)
        if ckpt:
            state = torch.load(ckpt, map_location="cpu", weights_only=False)
            vae.load_state_dict(state["model"], strict=True)
        vae.eval()
        for p in vae.parameters():
            p.requires_grad = False
        encoder = LatentDiscreteModel(vae, 512, args.latent_dim, timesteps=200, multitask=getattr(args, "multitask", False))
        return cls(args, encoder)

    @staticmethod
    def add_args(parser):
        add_shared_args(parser)
        parser.add_argument("--speech_decoder_ckpt", type=str, help="path to the speech decoder checkpoint")
        parser.add_argument("--use_cond", type=bool, default=False, help="use conditional diffusion")
        parser.add_argument("--multitask", type=bool, default=False)


@register_model_architecture("diff_discrete", "diff_discrete")
def base_architecture(args):
    args.attn_type = getattr(args, "attn_type", None)
    args.pos_enc_type = getattr(args, "pos_enc_type", "abs")
    args.classifier_guidance = getattr(args, "classifier_guidance", 1.0)
    args.latent_dim = getattr(args, "latent_dim", 16)
    args.multitask = getattr(args, "multitask", False)
