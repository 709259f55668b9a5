"""`speech_vae_decoder_loss` criterion (reference: fairseq/criterions/speech_vae_decoder_loss.py:14-129):
loss = 0.1 * LS-NLL(eps=0.1, ignore 0) / ntokens + 10 * mse + 1e-4 * kl  (:73-83)."""
from __future__ import annotations

import torch

from ..compat import FairseqCriterion, register_criterion


def label_smoothed_nll_loss(lprobs, target, epsilon, ignore_index=None, reduce=True):
    """Semantics of fairseq/criterions/label_smoothed_cross_entropy.py:34-51."""
    if target.dim() == lprobs.dim() - 1:
        target = target.unsqueeze(-1)
    nll = -lprobs.gather(dim=-1, index=target)
    smooth = -lprobs.sum(dim=-1, keepdim=True)
    if ignore_index is not None:
        pad = target.eq(ignore_index)
        nll, smooth = nll.masked_fill(pad, 0.0), smooth.masked_fill(pad, 0.0)
    else:
        nll, smooth = nll.squeeze(-1), smooth.squeeze(-1)
    if reduce:
        nll, smooth = nll.sum(), smooth.sum()
    e = epsilon / (lprobs.size(-1) - 1)
    return (1.0 - epsilon - e) * nll + e * smooth, nll


@register_criterion("speech_vae_decoder_loss")
class SpeechVAEDecoderLoss(FairseqCriterion):
    def __init__(self, task, label_smoothing=0.1):
        super().__init__(task)
        self.eps = label_smoothing
        self.padding_idx = 0

    def forward(self, model, sample, reduction="mean"):
        """:45-95 — trains on the REDUCED targets (:48-50), like the diffusion criterion."""
        tgt_feature, tgt_unit = sample["reduce_target"], sample["reduce_target_unit"]
        kwargs = dict(src_feature=sample["net_input"]["src_tokens"], src_lengths=sample["net_input"]["src_lengths"],
                      tgt_lengths=sample["reduce_target_lengths"], unk_token=self.task.tgt_dict.unk_index)
        mse_loss, lm_pred, kl_loss = model(tgt_feature, tgt_unit, **kwargs)
        lprobs = torch.log_softmax(lm_pred, dim=-1).view(-1, lm_pred.size(-1))
        target = tgt_unit.view(-1)
        tmask = target.ne(0)
        acc = torch.sum(lprobs.argmax(1).masked_select(tmask).eq(target.masked_select(tmask))) / torch.sum(tmask)
        loss_lm, nll = label_smoothed_nll_loss(lprobs, target, self.eps, ignore_index=0)
        ntokens = sample["ntokens"]
        loss = 0.1 * (loss_lm / ntokens) + 10 * mse_loss + 0.0001 * kl_loss
        sample_size = sample["nsentences"]
        logging_output = {"loss": loss.item(), "nll_loss": (nll / ntokens).item(), "mse_loss": mse_loss.item(),
                          "kl_loss": kl_loss.item(), "acc": acc.item(), "ntokens": ntokens,
                          "nsentences": sample["nsentences"], "sample_size": sample_size}
        return loss, sample_size, logging_output

    @staticmethod
    def logging_outputs_can_be_summed():
        return False
