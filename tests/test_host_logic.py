"""CPU: host-side logic of the package (module tree / state_dict parity, engine wiring, optimiser
hyper-parameters, derived metrics)."""
import math

import numpy as np
import torch

import continual_learning_b200 as clk
from continual_learning_b200.engine import UNetEngine
from oracle import metrics_ref
from oracle.unet_ref import layer_table, make_state_dict


def test_module_tree_matches_reference_state_dict_layout():
    m = clk.UNet(21)
    sd = make_state_dict(0)
    assert list(m.state_dict().keys()) == list(sd.keys())
    assert len(sd) == 136 and sum(p.numel() for p in m.parameters()) == 31044821
    assert m.load_state_dict(sd).missing_keys == []
    for k, v in m.state_dict().items():
        assert v.dtype == sd[k].dtype and v.shape == sd[k].shape


def test_engine_wiring_follows_layer_table():
    m = clk.UNet(7, in_dim=3, conv_dim=64)
    eng = UNetEngine(m)
    convs = [(p, ci, co) for p, kind, ci, co in layer_table(7) if kind == "conv3"]
    assert len(eng.units) == len(convs) == 18
    for u, (prefix, ci, co) in zip(eng.units, convs):
        assert u.cout == co, prefix
        assert (64 if u.stem else u.c0 + u.c1) == (64 if prefix == "enc1.0" else ci), prefix
        assert tuple(u.conv.weight.shape) == (co, ci, 3, 3)
    # concat units: skip first, upsampled second, equal halves (models/unet.py:83-87)
    for i in (10, 12, 14, 16):
        assert eng.units[i].c0 == eng.units[i].c1 > 0
    assert [co for (_, _, co) in eng.convT] == [512, 256, 128, 64]


def test_fused_adam_hyper_values_and_state_keys():
    p = torch.nn.Parameter(torch.zeros(4))
    opt = clk.FusedAdam([p], lr=1e-4, betas=(0.5, 0.99))
    lr, bc1, bc2s, gs = opt.hyper_values(3, grad_scale=0.5)
    assert lr == 1e-4 and gs == 0.5
    assert abs(bc1 - (1 - 0.5 ** 3)) < 1e-12 and abs(bc2s - math.sqrt(1 - 0.99 ** 3)) < 1e-12
    ref = torch.optim.Adam([torch.nn.Parameter(torch.zeros(4))], lr=1e-4, betas=(0.5, 0.99))
    assert set(opt.state_dict()["param_groups"][0]) >= {"lr", "betas", "eps", "weight_decay"}
    assert opt.state_dict()["param_groups"][0]["betas"] == ref.state_dict()["param_groups"][0]["betas"]


def test_derived_metrics_use_reference_float32_formulas():
    rng = np.random.Generator(np.random.PCG64(2))
    conf = torch.from_numpy(rng.integers(0, 1000, size=(22, 22), dtype=np.int64))
    conf[21] = 0
    conf[:, 21] = 0  # the extra 22nd class of trainer.py:188 is empty -> NaN -> filtered
    got = clk.metrics.metrics_from_matrix(conf)
    want = metrics_ref.derived_metrics(conf.float())
    assert [float(a) for a in got] == [float(b) for b in want]
    acc, tot, cor = clk.metrics.pixel_acc(torch.zeros(2, 3), None, 10.0, 4.0)
    assert (acc, tot, cor) == (100 * 4.0 / 16.0, 16.0, 4.0)


def test_fused_adam_step_count_survives_load_state_dict():
    """resume (trainer.py:98): the bias correction must continue from the checkpointed step, not restart at t = 1
    on warm moments; `TrainStep` re-reads the count whenever the optimiser state is replaced."""
    ps = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(3))]
    ref = torch.optim.Adam(ps, lr=1e-4, betas=(0.5, 0.99))
    for p in ps:
        p.grad = torch.ones_like(p)
    for _ in range(7):
        ref.step()
    qs = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(3))]
    opt = clk.FusedAdam(qs, lr=1e-4, betas=(0.5, 0.99))
    assert opt.group_step(qs) == 0.0 and opt.state_epoch == 0
    opt.load_state_dict(ref.state_dict())       # the reference's optimizer_state loads into the drop-in
    assert opt.state_epoch == 1 and opt.group_step(qs) == 7.0
    ts = clk.TrainStep(clk.UNet(21), opt)   # host-side object only: nothing is launched
    ts._sync_step_count()
    assert ts.step_count == 7
    assert opt.hyper_values(ts.step_count + 1)[1] == 1 - 0.5 ** 8
    opt.set_group_step(qs, 9)
    assert opt.group_step(qs) == 9.0 and opt.state[qs[0]]["step"] is opt.state[qs[1]]["step"]
    opt.state[qs[1]]["step"] = torch.tensor(3.0)
    try:
        opt.group_step(qs)
        raise AssertionError("disagreeing step counts must raise")
    except RuntimeError:
        pass
