"""Isolated import of the UNMODIFIED reference hot-path files (authoring container only).

TEST INFRASTRUCTURE ONLY (see oracle/diffnorm_oracle.py header).  ``import fairseq`` is impossible for
the reference as shipped, so this registers minimal stub modules for the handful of
``fairseq.*`` names ``latent_module.py`` imports and then loads the real files by path from
``/root/reference``.  Nothing is copied into this repo; on the GPU box ``/root/reference`` does not exist
and ``available()`` is False — only oracle/make_golden.py and the ``ref``-marked tests use this module.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
from torch import nn

# /root/reference in the authoring container; on the GPU box the untouched files of the path staged by `make -C oracle ref`
# under oracle/_ref/reference (git-ignored, travels with the snapshot; sha256 list in oracle/_ref/reference.sha256)
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference")
REF_ROOT = os.environ.get("DIFFNORM_REFERENCE", "/root/reference")
if not os.path.isfile(os.path.join(REF_ROOT, "fairseq", "models", "text_to_speech", "latent_module.py")):
    REF_ROOT = _STAGED
_TTS = os.path.join(REF_ROOT, "fairseq", "models", "text_to_speech")


def available() -> bool:
    return os.path.isfile(os.path.join(_TTS, "latent_module.py"))


def _load(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def load():
    """Returns a namespace with .latent_module, .distributions, .diffusion (generic lib package)."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")

    def pkg(name):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        return m

    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "fairseq" or k.startswith("fairseq.")}
    fairseq = pkg("fairseq")
    futils = types.ModuleType("fairseq.utils")

    def make_positions(tensor, padding_idx, onnx_trace=False):  # semantics of fairseq/utils.py:256-266
        valid = tensor.ne(padding_idx).int()
        return (torch.cumsum(valid, dim=1).type_as(valid) * valid).long() + padding_idx

    futils.make_positions = make_positions
    sys.modules["fairseq.utils"] = futils
    fairseq.utils = futils

    fmodules = pkg("fairseq.modules")
    sin = _load("fairseq.modules.sinusoidal_positional_embedding",
                os.path.join(REF_ROOT, "fairseq", "modules", "sinusoidal_positional_embedding.py"))

    def PositionalEmbedding(num_embeddings, embedding_dim, padding_idx, learned=False):
        assert not learned  # fairseq/modules/positional_embedding.py:29-34
        return sin.SinusoidalPositionalEmbedding(embedding_dim, padding_idx,
                                                 init_size=num_embeddings + padding_idx + 1)

    fmodules.PositionalEmbedding = PositionalEmbedding

    fmodels = pkg("fairseq.models")

    class FairseqEncoder(nn.Module):  # fairseq/models/fairseq_encoder.py:26-31
        def __init__(self, dictionary):
            super().__init__()
            self.dictionary = dictionary

    fmodels.FairseqEncoder = FairseqEncoder
    pkg("fairseq.criterions")
    lsce = types.ModuleType("fairseq.criterions.label_smoothed_cross_entropy")

    def label_smoothed_nll_loss(lprobs, target, epsilon, ignore_index=None, reduce=True):
        # semantics of fairseq/criterions/label_smoothed_cross_entropy.py:34-51
        if target.dim() == lprobs.dim() - 1:
            target = target.unsqueeze(-1)
        nll = -lprobs.gather(dim=-1, index=target)
        smooth = -lprobs.sum(dim=-1, keepdim=True)
        if ignore_index is not None:
            pad = target.eq(ignore_index)
            nll = nll.masked_fill(pad, 0.0)
            smooth = smooth.masked_fill(pad, 0.0)
        else:
            nll, smooth = nll.squeeze(-1), smooth.squeeze(-1)
        if reduce:
            nll, smooth = nll.sum(), smooth.sum()
        e = epsilon / (lprobs.size(-1) - 1)
        return (1.0 - epsilon - e) * nll + e * smooth, nll

    lsce.label_smoothed_nll_loss = label_smoothed_nll_loss
    sys.modules[lsce.__name__] = lsce
    if "sacrebleu" not in sys.modules:
        sys.modules["sacrebleu"] = types.ModuleType("sacrebleu")
    pkg("fairseq.models.text_to_speech")
    dist = _load("fairseq.models.text_to_speech.distributions", os.path.join(_TTS, "distributions.py"))
    lm = _load("fairseq.models.text_to_speech.latent_module", os.path.join(_TTS, "latent_module.py"))
    # generic diffusion lib: a package with only relative imports
    dpath = os.path.join(_TTS, "diffusion")
    spec = importlib.util.spec_from_file_location("refdiff", os.path.join(dpath, "__init__.py"),
                                                  submodule_search_locations=[dpath])
    refdiff = importlib.util.module_from_spec(spec)
    sys.modules["refdiff"] = refdiff
    spec.loader.exec_module(refdiff)
    ns = types.SimpleNamespace(latent_module=lm, distributions=dist, diffusion=refdiff)
    _cache["ns"] = ns
    # keep the stubs registered (the loaded modules reference them lazily) but remember the originals
    _cache["saved"] = saved
    return ns


def load_dataset_module():
    """The reference's training dataset, `fairseq/data/audio/repr_to_repr_unit_dataset.py`, and its real
    `fairseq/data/dictionary.py` + `fairseq/tokenizer.py`, loaded by path.  The dataset file imports a dozen fairseq.data
    names it never calls on this path (audio transforms, S2T dataset helpers): those are empty stubs; the data config is a
    stub exposing the three members the constructor reads (shuffle, get_feature_transforms, get_waveform_transforms)."""
    if "ds" in _cache:
        return _cache["ds"]
    load()   # registers the fairseq / fairseq.utils stubs
    fs = os.path.join(REF_ROOT, "fairseq")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _PathManager:
        open = staticmethod(open)

    mod("fairseq.file_chunker_utils", Chunker=object, find_offsets=lambda *a, **k: [])
    mod("fairseq.file_io", PathManager=_PathManager)
    _load("fairseq.tokenizer", os.path.join(fs, "tokenizer.py"))
    data = mod("fairseq.data", data_utils=types.ModuleType("fairseq.data.data_utils"))
    sys.modules["fairseq.data.data_utils"] = data.data_utils
    dic = _load("fairseq.data.dictionary", os.path.join(fs, "data", "dictionary.py"))
    data.Dictionary = dic.Dictionary
    data.FairseqDataset = torch.utils.data.Dataset
    data.ConcatDataset = torch.utils.data.ConcatDataset

    class _NoTransforms:
        @classmethod
        def from_config_dict(cls, cfg=None):
            return None

    class S2SDataConfig:  # fairseq/data/audio/data_cfg.py: only what repr_to_repr_unit_dataset.py:75-88 reads
        def __init__(self, shuffle=False):
            self.shuffle, self.use_audio_input = shuffle, False

        def get_feature_transforms(self, split, is_train):
            return None

        def get_waveform_transforms(self, split, is_train):
            return None

    mod("fairseq.data.audio")
    mod("fairseq.data.audio.audio_utils", get_features_or_waveform=None)
    mod("fairseq.data.audio.data_cfg", S2SDataConfig=S2SDataConfig)
    mod("fairseq.data.audio.speech_to_text_dataset", SpeechToTextDataset=object, SpeechToTextDatasetCreator=object,
        TextTargetMultitaskData=object, _collate_frames=None, _is_int_or_np_int=None)
    mod("fairseq.data.audio.feature_transforms", CompositeAudioFeatureTransform=_NoTransforms)
    mod("fairseq.data.audio.waveform_transforms", CompositeAudioWaveformTransform=_NoTransforms)
    mod("fairseq.data.audio.dataset_transforms", CompositeAudioDatasetTransform=_NoTransforms)
    mod("fairseq.data.audio.speech_to_speech_dataset", SpeechToSpeechDataset=object)
    ds = _load("fairseq.data.audio.repr_to_repr_unit_dataset",
               os.path.join(fs, "data", "audio", "repr_to_repr_unit_dataset.py"))
    _cache["ds"] = types.SimpleNamespace(module=ds, Dictionary=dic.Dictionary, S2SDataConfig=S2SDataConfig)
    return _cache["ds"]


def load_vocoder():
    """The reference's unit vocoder classes, loaded by path from the untouched files: hifigan.py (torch only),
    fastspeech2.py (for VariancePredictor; its other fairseq imports are stubbed with the semantics the class needs:
    FairseqDropout = nn.Dropout, LayerNorm = nn.LayerNorm) and codehifigan.py.  Returns the module of codehifigan.py."""
    if "voc" in _cache:
        return _cache["voc"]
    load()                                    # registers the fairseq.* stub packages
    fmodels, fmodules = sys.modules["fairseq.models"], sys.modules["fairseq.modules"]

    class FairseqDropout(nn.Dropout):          # fairseq/modules/fairseq_dropout.py: nn.Dropout with a module name
        def __init__(self, p, module_name=None):
            super().__init__(p)

    fmodules.FairseqDropout, fmodules.LayerNorm = FairseqDropout, nn.LayerNorm
    fmodules.MultiheadAttention = getattr(fmodules, "MultiheadAttention", nn.MultiheadAttention)
    fmodels.FairseqEncoderModel = getattr(fmodels, "FairseqEncoderModel", nn.Module)
    deco = lambda *a, **k: (lambda f: f)
    fmodels.register_model = getattr(fmodels, "register_model", deco)
    fmodels.register_model_architecture = getattr(fmodels, "register_model_architecture", deco)
    if "fairseq.data" not in sys.modules:
        d = types.ModuleType("fairseq.data")
        d.__path__ = []
        sys.modules["fairseq.data"] = d
    du = types.ModuleType("fairseq.data.data_utils")
    du.lengths_to_padding_mask = lambda lens: torch.arange(int(lens.max()))[None, :] >= lens[:, None]
    sys.modules["fairseq.data.data_utils"] = du
    for name, attr in (("hub_interface", "TTSHubInterface"), ("tacotron2", "Postnet")):
        m = types.ModuleType("fairseq.models.text_to_speech." + name)
        setattr(m, attr, type(attr, (nn.Module,), {}))
        sys.modules[m.__name__] = m
    _load("fairseq.models.text_to_speech.hifigan", os.path.join(_TTS, "hifigan.py"))
    _load("fairseq.models.text_to_speech.fastspeech2", os.path.join(_TTS, "fastspeech2.py"))
    voc = _load("fairseq.models.text_to_speech.codehifigan", os.path.join(_TTS, "codehifigan.py"))
    _cache["voc"] = voc
    return voc


def build_reference_model(latent_dim: int = 16, hid: int = 512, timesteps: int = 200, multitask: bool = False):
    """The reference's own modules: LatentDiscreteModel(vae, hid, z) (latent_module.py:1300)."""
    ns = load()
    vae = ns.latent_module.SpeechVAEEncoderDecoder(dim=768, latent_dim=latent_dim)
    ldm = ns.latent_module.LatentDiscreteModel(types.SimpleNamespace(encoder=vae), hid, latent_dim,
                                               timesteps=timesteps, multitask=multitask)
    return ldm.eval()


class ReplayNoise:
    """Context manager that makes torch.randn / randn_like return supplied tensors in call order, so the
    reference's CPU-drawn VAE noise (distributions.py:38), q_sample noise (latent_module.py:1409) and
    per-step draws (latent_module.py:1435) are deterministic and shared with the oracle."""

    def __init__(self, tensors):
        self.q = list(tensors)
        self.used = 0

    def __enter__(self):
        self._randn, self._randn_like = torch.randn, torch.randn_like

        def randn(*shape, **kw):
            if len(shape) == 1 and not isinstance(shape[0], int):
                shape = tuple(shape[0])
            if self.q and tuple(self.q[0].shape) == tuple(shape):
                self.used += 1
                return self.q.pop(0).clone()
            return self._randn(*shape, **kw)

        def randn_like(x, **kw):
            return self._randn_like(x, **kw)  # eta = 0: the per-step draw is discarded by the reference

        torch.randn, torch.randn_like = randn, randn_like
        return self

    def __exit__(self, *a):
        torch.randn, torch.randn_like = self._randn, self._randn_like
