"""`ddpm_discrete_loss` criterion (reference: fairseq/criterions/ddpm_discrete_loss.py:14-109)."""
from __future__ import annotations

from ..compat import FairseqCriterion, register_criterion


def _item(x):
    return x.item() if hasattr(x, "item") else x


@register_criterion("ddpm_discrete_loss")
class DDPMDiscreteLoss(FairseqCriterion):
    def __init__(self, task, label_smoothing=0.1):
        super().__init__(task)
        self.eps = label_smoothing
        self.padding_idx = 0  # the dataset pads unit labels with 0 (= <s>), repr_to_repr_unit_dataset.py:216-217

    def forward(self, model, sample, reduction="mean"):
        """:37-75 — reads the reduced targets and returns (loss, nsentences, logging dict)."""
        kwargs = dict(src_feature=sample["net_input"]["src_tokens"], src_lengths=sample["net_input"]["src_lengths"],
                      tgt_lengths=sample["reduce_target_lengths"], unk_token=self.task.tgt_dict.unk_index)
        loss_dict = model(sample["reduce_target"], sample["reduce_target_unit"], **kwargs)
        loss = loss_dict["total_loss"]
        sample_size = sample["nsentences"]
        logging_output = {
            "loss": _item(loss.data), "noise_loss": _item(loss_dict["noise_loss"].data),
            "nll_loss": _item(loss_dict["nll_loss"].data), "mse_loss": _item(loss_dict["recon_mse_loss"].data),
            "acc": _item(loss_dict["acc"].data), "ntokens": sample["ntokens"], "nsentences": sample["nsentences"],
            "sample_size": sample_size,
        }
        return loss, sample_size, logging_output

    @staticmethod
    def reduce_metrics(logging_outputs):
        """:77-95 — sample-size-weighted means over the workers (logging outputs cannot be summed)."""
        ns = [lo.get("sample_size", 0) for lo in logging_outputs]
        ntot = sum(ns)
        ws = [n / (ntot + 1e-8) for n in ns]
        red = {k: sum(lo.get(k, 0) * w for lo, w in zip(logging_outputs, ws)) for k in ("loss", "noise_loss", "mse_loss", "nll_loss", "acc")}
        red["sample_size"] = ntot
        return red

    @staticmethod
    def logging_outputs_can_be_summed():
        return False
