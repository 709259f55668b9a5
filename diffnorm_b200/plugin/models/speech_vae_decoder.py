"""`speech_vae_decoder` model + arch (reference: fairseq/models/text_to_speech/speech_vae_decoder.py:25-136)."""
from __future__ import annotations

import torch

from ..compat import FairseqEncoderModel, lengths_to_mask, register_model, register_model_architecture
from ..latent_module import SpeechVAEEncoderDecoder


def add_shared_args(parser):
    """Flags both model plugins accept (speech_vae_decoder.py:69-96, diff_discrete.py:88-130).  The reference also
    pulls in the dead DiffusionTransformerModel.add_args (diff_transformer.py:190); only the flags its recipes
    pass (scripts/vae/train.sh, scripts/diffusion/train.sh) are kept."""
    parser.add_argument("--input-feat-per-channel", default=80)
    parser.add_argument("--depthwise-conv-kernel-size", default=31)
    parser.add_argument("--input-channels", default=1)
    parser.add_argument("--attn-type", default=None)
    parser.add_argument("--pos-enc-type", default="abs")
    parser.add_argument("--classifier_guidance", type=float, default=1.0)
    parser.add_argument("--latent_dim", type=int, default=16)


@register_model("speech_vae_decoder")
class SpeechVAEDecoder(FairseqEncoderModel):
    def __init__(self, args, encoder):
        super().__init__(encoder)
        self.args = args

    def forward(self, target_feature, target_unit, **model_kwargs):
        tgt_mask = lengths_to_mask(model_kwargs["tgt_lengths"])
        return self.encoder(target_feature, target_unit, tgt_mask)  # (mse_loss, lm_pred, kl_loss), :35-44

    def get_normalized_probs(self, net_output, log_probs, sample=None):
        logits = net_output[0]
        return torch.log_softmax(logits, dim=-1) if log_probs else torch.softmax(logits, dim=-1)

    @classmethod
    def build_model(cls, args, task):
        return cls(args, SpeechVAEEncoderDecoder(dim=768, latent_dim=args.latent_dim))

    @staticmethod
    def add_args(parser):
        add_shared_args(parser)


@register_model_architecture("speech_vae_decoder", "speech_vae_decoder")
def base_architecture(args):
    args.attn_type = getattr(args, "attn_type", None)
    args.pos_enc_type = getattr(args, "pos_enc_type", "abs")
    args.classifier_guidance = getattr(args, "classifier_guidance", 1.0)
    args.latent_dim = getattr(args, "latent_dim", 16)
