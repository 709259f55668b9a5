"""`speech_diffusion_discrete` task (reference: fairseq/tasks/speech_diffusion_discrete_task.py:33-261): the task a
diffusion checkpoint names, hence the one `load_model_ensemble_and_task` sets up before the normalization driver
calls `model.encoder.ddim_sample` (diff_norm_synthesis.py:185-204).  Same args / dictionary as `speech_decoder`."""
from __future__ import annotations

from ..compat import register_task
from .speech_decoder_task import _UnitTask


@register_task("speech_diffusion_discrete")
class SpeechDiffusionDiscreteTask(_UnitTask):
    pass
