"""fairseq plugin package for the DiffNorm normalization path (usable as ``--user-dir diffnorm_b200/plugin``).

Registers, under the reference's names:
  tasks       speech_decoder, speech_diffusion_discrete
  models      speech_vae_decoder (+arch), diff_discrete (+arch)
  criterions  speech_vae_decoder_loss, ddpm_discrete_loss
fairseq's ``import_user_module`` imports this package and then its ``tasks`` / ``models`` sub-packages
(fairseq/utils.py:464-509); criterions are imported from here because fairseq does not scan for them.
"""
from . import compat  # noqa: F401
from . import models, tasks, criterions  # noqa: F401,E402
