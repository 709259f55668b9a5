from .speech_vae_decoder import SpeechVAEDecoder  # noqa: F401
from .diff_discrete import DiffDiscreteModel  # noqa: F401
