"""`speech_decoder` task (reference: fairseq/tasks/speech_decoder_task.py:33-258): VAE training/eval task.
Argument names and defaults follow :36-116; the unit dictionary follows :137-141."""
from __future__ import annotations

import torch

from ..compat import Dictionary, LegacyFairseqTask, register_task


def add_task_args(parser):
    parser.add_argument("data", help="manifest root")
    parser.add_argument("--config-yaml", type=str, default="config.yaml")
    parser.add_argument("--max-source-positions", default=6000, type=int, metavar="N")
    parser.add_argument("--max-target-positions", default=1024, type=int, metavar="N")
    parser.add_argument("--target-is-code", action="store_true")
    parser.add_argument("--target-code-size", type=int, default=None, help="# discrete units")
    parser.add_argument("--save-audio", action="store_true")
    parser.add_argument("--n-frames-per-step", type=int, default=1)
    parser.add_argument("--eval-inference", action="store_true")
    parser.add_argument("--eval-args", type=str, default="{}")
    parser.add_argument("--eos-prob-threshold", type=float, default=0.5)
    parser.add_argument("--mcd-normalize-type", type=str, default="targ", choices=["targ", "pred", "path"])
    parser.add_argument("--vocoder", type=str, default="griffin_lim", choices=["griffin_lim", "hifigan", "code_hifigan"])
    parser.add_argument("--spec-bwd-max-iter", type=int, default=8)
    parser.add_argument("--infer-target-lang", type=str, default="")
    parser.add_argument("--dummy-config", type=str, default=None)
    parser.add_argument("--vocoder-config", type=str, default=None)
    parser.add_argument("--tgt-feat-dir", type=str, default=None)
    parser.add_argument("--src-feat-dir", type=str, default=None)


class _UnitTask(LegacyFairseqTask):
    def __init__(self, args, tgt_dict):
        super().__init__(args)
        self.tgt_dict = tgt_dict

    @staticmethod
    def add_args(parser):
        add_task_args(parser)

    @classmethod
    def setup_task(cls, args, **kwargs):
        tgt_dict = None
        if getattr(args, "target_is_code", False):
            assert args.target_code_size is not None
            tgt_dict = Dictionary()
            for i in range(args.target_code_size):
                tgt_dict.add_symbol(str(i))
        if getattr(args, "train_subset", None) is not None:
            if not all(s.startswith("train") for s in args.train_subset.split(",")):
                raise ValueError('Train splits should be named like "train*".')
        assert getattr(args, "n_frames_per_step", 1) >= 1
        return cls(args, tgt_dict)

    @property
    def target_dictionary(self):
        return self.tgt_dict

    @property
    def source_dictionary(self):
        return None

    def max_positions(self):
        return getattr(self.args, "max_source_positions", 6000), getattr(self.args, "max_target_positions", 1024)

    def load_dataset(self, split, epoch=1, combine=False, **kwargs):
        from ...data import ReprToReprUnitDataset
        self.datasets[split] = ReprToReprUnitDataset.from_manifest(self.args, split, self.tgt_dict)

    def train_step(self, sample, model, criterion, optimizer, update_num, ignore_grad=False):
        model.train()
        model.set_num_updates(update_num)
        loss, sample_size, logging_output = criterion(model, sample)
        if ignore_grad:
            loss *= 0
        optimizer.backward(loss)
        return loss, sample_size, logging_output

    def valid_step(self, sample, model, criterion):
        model.eval()
        with torch.no_grad():
            return criterion(model, sample)


@register_task("speech_decoder")
class SpeechDecoderTask(_UnitTask):
    pass
