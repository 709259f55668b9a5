from .speech_vae_decoder_loss import SpeechVAEDecoderLoss  # noqa: F401
from .ddpm_discrete_loss import DDPMDiscreteLoss  # noqa: F401
