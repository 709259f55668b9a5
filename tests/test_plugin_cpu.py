"""The fairseq plugin surface: registry names, dictionary, state_dict key compatibility,
checkpoint round trip, error behaviour — all on CPU (compute needs a GPU and raises without one)."""
import argparse

import pytest
import torch

from diffnorm_b200.plugin import compat
from oracle import diffnorm_oracle as O


def _args(**kw):
    d = dict(task="speech_diffusion_discrete", arch="diff_discrete", target_is_code=True, target_code_size=1000,
             latent_dim=16, criterion="ddpm_discrete_loss")
    d.update(kw)
    return argparse.Namespace(**d)


@pytest.mark.skipif(compat.HAVE_FAIRSEQ, reason="shim registries only exist without fairseq")
def test_registry_names_and_duplicates():
    assert {"speech_decoder", "speech_diffusion_discrete"} <= set(compat.TASK_REGISTRY)
    assert {"speech_vae_decoder", "diff_discrete"} <= set(compat.MODEL_REGISTRY)
    assert {"speech_vae_decoder", "diff_discrete"} <= set(compat.ARCH_MODEL_REGISTRY)
    assert {"speech_vae_decoder_loss", "ddpm_discrete_loss"} <= set(compat.CRITERION_REGISTRY)
    with pytest.raises(ValueError, match="duplicate task"):
        compat.register_task("speech_decoder")(compat.TASK_REGISTRY["speech_decoder"])
    with pytest.raises(ValueError, match="duplicate model"):
        compat.register_model("diff_discrete")(compat.MODEL_REGISTRY["diff_discrete"])
    with pytest.raises(ValueError, match="unknown model type"):
        compat.register_model_architecture("nope", "nope")(lambda a: None)
    with pytest.raises(ValueError):
        compat.setup_task(_args(task="missing"))


def test_dictionary_and_units_offset():
    task = compat.setup_task(_args())
    d = task.target_dictionary
    assert len(d) == 1004 and d.index("0") == 4 and d.index("999") == 1003  # unit k <-> index k + 4 (LM:1451)
    assert (d.bos(), d.pad(), d.eos(), d.unk()) == (0, 1, 2, 3)


@pytest.mark.parametrize("z", [16, 128])
def test_state_dict_keys_match_reference_layout(z):
    task = compat.setup_task(_args(latent_dim=z))
    model = task.build_model(_args(latent_dim=z), from_checkpoint=True)
    sd = model.state_dict()
    ref = O.init_state_dict(O.Arch(latent_dim=z), seed=3)
    assert set(sd) == {"encoder." + k for k in ref}
    if z == 16:   # ORDER too (parameter registration order; golden from the live reference, oracle/make_golden.py make_keys)
        import os
        want = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "state_dict_keys_z16.txt")).read().split()
        assert list(sd) == ["encoder." + k for k in want]
        assert [n for n, _ in model.named_parameters()] == ["encoder." + k for k in want if not k.endswith("_float_tensor")]
    assert all(tuple(sd["encoder." + k].shape) == tuple(v.shape) for k, v in ref.items())
    model.load_state_dict({"encoder." + k: v for k, v in ref.items()}, strict=True)
    assert torch.equal(model.encoder.model.final_proj.weight, ref["model.final_proj.weight"])
    # VAE-only checkpoint layout (speech_vae_decoder): encoder.encoder_wave.* etc.
    vae = compat.ARCH_MODEL_REGISTRY["speech_vae_decoder"].build_model(_args(latent_dim=z), None)
    assert set(vae.state_dict()) == {"encoder." + k[len("speech_decoder."):] for k in ref if k.startswith("speech_decoder.")}
    frozen = [p.requires_grad for p in model.encoder.speech_decoder.parameters()]
    assert not any(frozen)  # diff_discrete.py:77-79


def test_compute_requires_cuda_and_training_not_silently_faked():
    task = compat.setup_task(_args())
    model = task.build_model(_args(), from_checkpoint=True)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            model.encoder.ddim_sample(torch.zeros(1, 8, 768), input_mask=torch.ones(1, 8, dtype=torch.bool), start_step=5)
    if not torch.cuda.is_available():   # the training step has no CPU path either: it raises, it does not fake a loss
        with pytest.raises(RuntimeError, match="CUDA"):
            model.encoder(torch.zeros(1, 8, 768), torch.zeros(1, 8, dtype=torch.long), tgt_mask=torch.ones(1, 8, dtype=torch.bool))
    red = type(task.build_criterion(_args())).reduce_metrics([{"loss": 1.0, "sample_size": 1}, {"loss": 3.0, "sample_size": 3}])
    assert abs(red["loss"] - 2.5) < 1e-6 and red["sample_size"] == 4   # ddpm_discrete_loss.py:77-95 weighting
    crit = task.build_criterion(_args())
    assert type(crit).__name__ == "DDPMDiscreteLoss" and crit.logging_outputs_can_be_summed() is False


def test_scheduler_has_the_reference_signature_and_accessors():
    from diffnorm_b200.plugin.latent_module import DDPMScheduler
    s = DDPMScheduler(200, 1.0)          # (timesteps, scale) like LM:1242
    o = O.Schedule(200)
    t = torch.tensor([0, 99, 199])
    assert s.scale == 1.0 and s.num_timesteps == 200
    got = s.get_sqrt_alpha_cum(t, (3, 4, 5))
    assert got.shape == (3, 4, 5) and got.dtype == torch.float32
    assert torch.equal(got[:, 0, 0], torch.from_numpy(o.sqrt_alphas_cumprod)[t].float())
    assert torch.equal(s.get_alpha_prev_cum(t, (3,)), torch.from_numpy(o.alphas_cumprod_prev)[t].float())
    assert torch.equal(s.get_beta(t, (3,)), torch.from_numpy(o.betas)[t].float())
    snr = s.get_snr(t)
    assert torch.allclose(snr, torch.from_numpy(o.alphas_cumprod / (1 - o.alphas_cumprod))[t].float(), rtol=1e-5)


def test_mask_must_be_prefix():
    from diffnorm_b200.plugin.latent_module import _mask_to_lengths
    m = torch.tensor([[True, True, False], [True, False, True]])
    with pytest.raises(ValueError):
        _mask_to_lengths(m)
    assert _mask_to_lengths(torch.tensor([[True, True, False], [True, False, False]])).tolist() == [2, 1]
