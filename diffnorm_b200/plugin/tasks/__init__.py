from .speech_decoder_task import SpeechDecoderTask  # noqa: F401
from .speech_diffusion_discrete_task import SpeechDiffusionDiscreteTask  # noqa: F401
