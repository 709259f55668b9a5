"""The descriptor table of the one-launch weight re-packing (diffnorm_b200.repack) against diffnorm_b200.packing, on CPU:
the table is recorded next to every pack call of the training step, executed with the plain-torch reference executor after
the master weights change, and must reproduce bit for bit what packing.* builds from the new weights."""
import torch

from diffnorm_b200.config import DiffNormConfig
from diffnorm_b200.packing import rup
from diffnorm_b200.plugin.latent_module import Model
from diffnorm_b200.repack import PackTable, plan_tensors
from diffnorm_b200.train import DenoiserTrainer


def small_trainer(hid=128, depth=2, stacks=2, layers=3, z=16):
    torch.manual_seed(0)
    cfg = DiffNormConfig(latent_dim=z, hid=hid, depth=depth, wn_stacks=stacks, wn_layers=layers)
    model = Model(hid, z, cfg)
    tr = object.__new__(DenoiserTrainer)          # packing only: no device state
    tr.cfg, tr.P, tr.dev = cfg, dict(model.named_parameters()), torch.device("cpu")
    tr.inner = DiffNormConfig.ff_inner(hid)
    tr.ip = rup(tr.inner, 128)
    tr.zp, tr.zn = rup(z, 64), rup(z, 16)
    tr.cond_names = DenoiserTrainer.cond_layer_names(cfg)
    return tr


def test_recorded_table_reproduces_packing_after_weight_update():
    tr = small_trainer()
    rec = PackTable()
    plans = tr._pack(rec)
    named = plan_tensors(plans)
    assert len(rec.ops) > 50 and len(named) > 40
    before = {n: t.clone() for n, t in named}
    with torch.no_grad():
        for p in tr.P.values():                   # an "optimizer step" in place
            p.add_(torch.randn_like(p) * 0.05)
    rec.run_reference()
    fresh = dict(plan_tensors(tr._pack()))
    changed = 0
    for n, t in named:
        assert t.dtype == fresh[n].dtype and t.shape == fresh[n].shape, n
        assert torch.equal(t, fresh[n]), f"{n}: table and packing.* disagree"
        changed += int(not torch.equal(t, before[n]))
    assert changed == len(named)                  # every packed tensor is covered by the table


def test_table_tiles_are_a_partition():
    tr = small_trainer(depth=1, stacks=1, layers=2)
    rec = PackTable()
    tr._pack(rec)
    rows, total = rec._rows()
    nxt = 0
    for o, tile0, tiles_c in rows:
        assert tile0 == nxt and tiles_c == -(-o["cols"] // 64)
        nxt += -(-o["rows"] // 64) * tiles_c
    assert nxt == total


def test_vae_recorded_table_reproduces_packing_after_weight_update():
    """Same check for the VAE training step's packings (WaveNet encoder / decoder blocks with channel padding, the
    768-wide decoder transformer, the per-layer gamma copies)."""
    from diffnorm_b200.plugin.latent_module import SpeechVAEEncoderDecoder
    from diffnorm_b200.train import FrozenDecoderTrain
    from diffnorm_b200.train_vae import VaeTrainer
    torch.manual_seed(1)
    vae = SpeechVAEEncoderDecoder(192, 16)        # narrow features keep the CPU test small; same code paths
    cfg = vae.cfg
    cfg.vae_depth = 1
    tr = object.__new__(VaeTrainer)
    tr.cfg, tr.P, tr.dev = cfg, dict(vae.named_parameters()), torch.device("cpu")
    tr.G, tr.S = cfg.vae_layers, cfg.vae_stacks
    tr.dec = FrozenDecoderTrain(None, cfg, tr.dev, None)
    rec = PackTable()
    enc_blocks = tr._pack(rec)

    def tree():
        return plan_tensors([enc_blocks, tr.dec.blocks, tr.dec.layers, tr.dec.pred, tr.dec.pred_T, tr.dec.lm, tr.dec.lm_T,
                             tr.dec.pred_gamma])
    named = tree()
    before = {n: t.clone() for n, t in named}
    with torch.no_grad():
        for p in tr.P.values():
            p.add_(torch.randn_like(p) * 0.05)
    rec.run_reference()
    enc_blocks = tr._pack()
    fresh = dict(tree())
    assert len(named) > 60
    for n, t in named:
        assert torch.equal(t, fresh[n]), f"{n}: table and packing.* disagree"
        assert not torch.equal(t, before[n]), f"{n} is not covered by the table"
