"""fairseq plugin plumbing.  When fairseq is importable the plugin registers into fairseq's own registries
(the path fairseq takes for ``--user-dir``, fairseq/utils.py:464-509); otherwise a local shim with the same
decorator semantics (duplicate-name ValueError etc., fairseq/models/__init__.py:129-205,
fairseq/tasks/__init__.py:49-101, fairseq/registry.py:17-100) is used so the surface is testable stand-alone.
"""
from __future__ import annotations

import torch
from torch import nn

try:  # pragma: no cover - exercised only where fairseq is installed
    from fairseq.criterions import FairseqCriterion, register_criterion  # type: ignore
    from fairseq.data import Dictionary  # type: ignore
    from fairseq.models import (FairseqEncoder, FairseqEncoderModel, register_model,  # type: ignore
                                register_model_architecture)
    from fairseq.tasks import LegacyFairseqTask, register_task  # type: ignore
    HAVE_FAIRSEQ = True
    TASK_REGISTRY = MODEL_REGISTRY = ARCH_MODEL_REGISTRY = ARCH_CONFIG_REGISTRY = CRITERION_REGISTRY = None
except Exception:  # noqa: BLE001 - any import failure (missing hydra/omegaconf/...) selects the shim
    HAVE_FAIRSEQ = False
    TASK_REGISTRY, MODEL_REGISTRY, ARCH_MODEL_REGISTRY, ARCH_CONFIG_REGISTRY, CRITERION_REGISTRY = {}, {}, {}, {}, {}

    class Dictionary:
        """Subset of fairseq.data.Dictionary: specials <s>,<pad>,</s>,<unk> at 0..3 (dictionary.py:34-37)."""

        def __init__(self, bos="<s>", pad="<pad>", eos="</s>", unk="<unk>"):
            self.symbols, self.indices = [], {}
            self.bos_index = self.add_symbol(bos)
            self.pad_index = self.add_symbol(pad)
            self.eos_index = self.add_symbol(eos)
            self.unk_index = self.add_symbol(unk)
            self.nspecial = len(self.symbols)

        def add_symbol(self, word, n=1):
            if word in self.indices:
                return self.indices[word]
            self.indices[word] = len(self.symbols)
            self.symbols.append(word)
            return self.indices[word]

        def index(self, sym):
            return self.indices.get(sym, self.unk_index)

        def __len__(self):
            return len(self.symbols)

        def __getitem__(self, i):
            return self.symbols[i] if i < len(self.symbols) else "<unk>"

        def bos(self):
            return self.bos_index

        def pad(self):
            return self.pad_index

        def eos(self):
            return self.eos_index

        def unk(self):
            return self.unk_index

    class FairseqEncoder(nn.Module):  # fairseq/models/fairseq_encoder.py:26-31
        def __init__(self, dictionary):
            super().__init__()
            self.dictionary = dictionary

        def max_positions(self):
            return 1e6

    class BaseFairseqModel(nn.Module):
        @classmethod
        def add_args(cls, parser):
            pass

        @classmethod
        def build_model(cls, args, task):
            raise NotImplementedError

        def set_num_updates(self, num_updates):
            pass

    class FairseqEncoderModel(BaseFairseqModel):  # fairseq/models/fairseq_model.py
        def __init__(self, encoder):
            super().__init__()
            self.encoder = encoder

        def max_positions(self):
            return self.encoder.max_positions()

    class FairseqCriterion(nn.Module):  # fairseq/criterions/fairseq_criterion.py:15-22
        def __init__(self, task):
            super().__init__()
            self.task = task
            if hasattr(task, "target_dictionary"):
                d = task.target_dictionary
                self.padding_idx = d.pad() if d is not None else -100

        @classmethod
        def build_criterion(cls, cfg, task):
            return cls(task)

        @staticmethod
        def logging_outputs_can_be_summed():
            return False

    class LegacyFairseqTask:  # fairseq/tasks/fairseq_task.py (the members the plugins rely on)
        def __init__(self, args):
            self.args = args
            self.datasets = {}

        @staticmethod
        def add_args(parser):
            pass

        @classmethod
        def setup_task(cls, args, **kwargs):
            return cls(args, **kwargs)

        def build_model(self, args, from_checkpoint=False):
            return build_model(args, self, from_checkpoint)

        def build_criterion(self, args):
            return CRITERION_REGISTRY[args.criterion].build_criterion(args, self)

        def dataset(self, split):
            return self.datasets[split]

    def register_task(name, dataclass=None):
        def deco(cls):
            if name in TASK_REGISTRY:
                raise ValueError("Cannot register duplicate task ({})".format(name))
            if not issubclass(cls, LegacyFairseqTask):
                raise ValueError("Task ({}: {}) must extend FairseqTask".format(name, cls.__name__))
            TASK_REGISTRY[name] = cls
            return cls
        return deco

    def register_model(name, dataclass=None):
        def deco(cls):
            if name in MODEL_REGISTRY:
                raise ValueError("Cannot register duplicate model ({})".format(name))
            if not issubclass(cls, BaseFairseqModel):
                raise ValueError("Model ({}: {}) must extend BaseFairseqModel".format(name, cls.__name__))
            MODEL_REGISTRY[name] = cls
            return cls
        return deco

    def register_model_architecture(model_name, arch_name):
        def deco(fn):
            if model_name not in MODEL_REGISTRY:
                raise ValueError("Cannot register model architecture for unknown model type ({})".format(model_name))
            if arch_name in ARCH_MODEL_REGISTRY:
                raise ValueError("Cannot register duplicate model architecture ({})".format(arch_name))
            if not callable(fn):
                raise ValueError("Model architecture must be callable ({})".format(arch_name))
            ARCH_MODEL_REGISTRY[arch_name] = MODEL_REGISTRY[model_name]
            ARCH_CONFIG_REGISTRY[arch_name] = fn
            return fn
        return deco

    def register_criterion(name, dataclass=None):
        def deco(cls):
            if name in CRITERION_REGISTRY:
                raise ValueError("Cannot register duplicate criterion ({})".format(name))
            CRITERION_REGISTRY[name] = cls
            return cls
        return deco

    def build_model(args, task, from_checkpoint=False):
        """fairseq/models/__init__.py:56-106 for legacy (argparse Namespace) configs."""
        arch = getattr(args, "arch", None)
        if arch not in ARCH_MODEL_REGISTRY:
            raise ValueError(f"Could not infer model type from arch={arch!r}. Available: {sorted(ARCH_MODEL_REGISTRY)}")
        ARCH_CONFIG_REGISTRY[arch](args)
        return ARCH_MODEL_REGISTRY[arch].build_model(args, task)

    def setup_task(args, **kwargs):
        name = getattr(args, "task", None)
        if name not in TASK_REGISTRY:
            raise ValueError(f"Could not infer task type from task={name!r}. Available: {sorted(TASK_REGISTRY)}")
        return TASK_REGISTRY[name].setup_task(args, **kwargs)


def lengths_to_mask(lens: torch.Tensor) -> torch.Tensor:
    """fairseq/data/data_utils.py:542-552 (True = valid frame)."""
    max_len = int(torch.max(lens).item())
    return torch.arange(max_len, device=lens.device).view(1, max_len) < lens.view(-1, 1)
