"""CPU-side checks of the host mirror: class surface, state_dict keys/shapes, RNG consumption,
error behaviour -- everything of SURVEY 8(b) that does not need a kernel launch."""
import numpy as np
import pytest
import torch

import pytorch_glow_b200 as G
from pytorch_glow_b200 import _C
from pytorch_glow_b200.hps import Hps, make_hps
from oracle import glow_oracle as O


def test_state_dict_matches_reference_keys_and_shapes(golden_glow):
    for perm, coup in (("invconv", "affine"), ("reverse", "additive")):
        ref = golden_glow.sd("%s_%s/sd/" % (perm, coup))
        hps = make_hps((16, 16, 3), K=2, L=2, hidden_channels=16, coupling=coup, permutation=perm, batch=4)
        mine = G.Glow(hps).state_dict()
        assert list(mine.keys()) == list(ref.keys())
        for k in ref:
            assert tuple(mine[k].shape) == tuple(ref[k].shape), k
        G.Glow(hps).load_state_dict(ref)        # a reference snapshot loads as-is


def test_celeba_profile_layout():
    hps = make_hps()        # profile/celeba.json shapes
    with torch.device("meta"):
        glow = G.Glow(hps)
    layers = glow.flow.layers
    assert len(layers) == 101 and glow.flow.output_shapes[-1] == [-1, 48, 8, 8]
    n_params = sum(p.numel() for n, p in glow.named_parameters() if n != "h_top")
    assert n_params == 44052720
    assert tuple(glow.h_top.shape) == (16, 96, 8, 8) and glow.batch_h_top == 16
    kinds = [type(m).__name__ for m in layers]
    assert kinds[0] == "Squeeze2d" and kinds[33] == "Split2d" and kinds[34] == "Squeeze2d" and kinds[-1] == "FlowStep"
    _, shapes = O.flow_layout((64, 64, 3), 32, 3)
    assert shapes == glow.flow.output_shapes


def test_h_top_follows_device_count():
    hps = make_hps(batch=16, devices=("cuda:0", "cuda:1"))
    with torch.device("meta"):
        assert G.Glow(hps).h_top.shape[0] == 8      # num_batch_train // num_device (model.py:350-353)


def test_numpy_rng_consumption_matches_reference(golden_layers):
    np.random.seed(3)
    assert torch.equal(G.Invertible1x1Conv(5).weight.data, golden_layers.t("invconv/init_seed3_c5"))
    np.random.seed(0)
    pm = G.Permutation2d(6, shuffle=True)
    assert pm.indices.tolist() == [0, 3, 4, 2, 5, 1] and pm.indices_inverse.tolist() == [0, 5, 3, 1, 2, 4]
    assert G.Permutation2d(6).indices.tolist() == [5, 4, 3, 2, 1, 0]
    assert pm.indices.dtype == np.int64


def test_constructor_errors_and_class_surface():
    with pytest.raises(AssertionError):
        G.FlowStep(8, 16, permutation="bogus")
    with pytest.raises(AssertionError):
        G.FlowStep(8, 16, coupling="bogus")
    with pytest.raises(AssertionError):
        G.FlowModel((16, 16, 2), 16, 1, 1)        # C must be 1 or 3 (model.py:233)
    assert G.FlowStep.flow_permutation_list == ['invconv', 'reverse', 'shuffle']
    assert G.FlowStep.flow_coupling_list == ['additive', 'affine']
    assert G.FlowNet is G.FlowModel and G.InvertibleConv1x1 is G.Invertible1x1Conv and G.SqueezeLayer is G.Squeeze2d
    z = G.Conv2dZeros(16, 5)
    assert tuple(z.weight.shape[:2]) == (5, 16) and float(z.weight.abs().sum()) == 0     # test_module.py:47
    lz = G.LinearZeros(16, 16)
    assert torch.equal(lz(torch.rand(16)), torch.zeros(16))                               # test_module.py:29
    assert G.Conv2d.get_padding('SAME', (3, 3), 1) == (1, 1) and G.Conv2d.get_padding('VALID', 3, 1) == (0, 0)
    assert abs(G.GaussianDiag.log_2pi - float(np.log(2 * np.pi))) < 1e-12


def test_no_cpu_fallback():
    fs = G.FlowStep(8, 16)
    with pytest.raises(_C.GlowkError):
        fs(torch.zeros(1, 8, 4, 4), None)
    with pytest.raises(_C.GlowkError):
        G.Squeeze2d()(torch.zeros(1, 3, 4, 4))
    with pytest.raises(AssertionError):
        fs(torch.zeros(1, 7, 4, 4))               # odd channels (model.py:169)


def test_set_actnorm_inited():
    hps = make_hps((16, 16, 3), K=1, L=2, hidden_channels=16, batch=2)
    glow = G.Glow(hps)
    acts = [m for m in glow.modules() if isinstance(m, G.ActNorm)]
    assert len(acts) == 6 and not any(m.bias_inited for m in acts)
    glow.set_actnorm_inited()
    assert all(m.bias_inited and m.logs_inited for m in acts)
    assert acts[0].needs_init is False
    glow.set_actnorm_inited(False)
    assert acts[0].needs_init is True
    glow.eval()
    assert acts[0].needs_init is False          # eval never initialises (module.py:93-94)


def test_lu_parameterisation_imports_dense_weight():
    torch.manual_seed(0)
    w = torch.randn(12, 12)
    np.random.seed(0)
    lu = G.Invertible1x1Conv(12, lu_decomposition=True)
    assert set(lu.state_dict().keys()) == {"p", "sign_s", "l", "u", "log_s"}
    lu.load_state_dict({"weight": w})
    w2, logabs = O.lu_assemble(lu.p, lu.l.data, lu.u.data, lu.sign_s, lu.log_s.data)
    assert float((w2 - w).abs().max()) < 1e-4
    assert abs(float(logabs) - float(torch.log(torch.abs(torch.det(w))))) < 1e-4


def test_hps_container():
    h = Hps({"a": {"b": 3}, "c": [1, 2]})
    assert h.a.b == 3 and h["a"]["b"] == 3
    h.a.b = 4
    import json
    assert json.loads(json.dumps(h))["a"]["b"] == 4


def test_rows_path_level_dispatch():
    """Which layers the pixel-major path takes (rows_path._prefix_len): every level up to 96 channels, 1x1-conv
    levels up to 384 channels (fp32 GEMM mix), and only whole levels ending in a Split2d before a wide permutation level."""
    from pytorch_glow_b200 import rows_path
    with torch.device("meta"):
        celeba = G.FlowModel((64, 64, 3), 512, K=4, L=3)
        hq = G.FlowModel((256, 256, 3), 512, K=2, L=6)                          # 12 ... 384 channels, invconv
        hq_perm = G.FlowModel((64, 64, 3), 32, K=2, L=6, permutation="reverse")  # wide permutation levels
        too_wide = G.FlowModel((128, 128, 3), 32, K=1, L=7)                      # level 7: 768 channels
    assert _C.lib().glowk_rows_max_channels() == 96 and _C.lib().glowk_rows_max_channels_wide() == 384
    assert rows_path._prefix_len(celeba) == len(celeba.layers)
    assert rows_path._prefix_len(hq) == len(hq.layers)
    k = rows_path._prefix_len(hq_perm)
    assert 0 < k < len(hq_perm.layers)
    assert isinstance(hq_perm.layers[k - 1], G.Split2d) and isinstance(hq_perm.layers[k], G.Squeeze2d)
    assert hq_perm.layers[k - 2].in_channels == 96 and hq_perm.layers[k + 1].in_channels == 192
    k7 = rows_path._prefix_len(too_wide)
    assert isinstance(too_wide.layers[k7 - 1], G.Split2d) and too_wide.layers[k7 + 1].in_channels == 768


def test_standalone_modules_refuse_silent_graph_cuts():
    """stand-alone Conv2d / f() ran on detached weights and returned tensors without grad_fn.  They now
    raise when a gradient is wanted; Conv2dZeros (learn_top, network/model.py:371-373) stays differentiable."""
    import pytest
    import pytorch_glow_b200 as G
    from pytorch_glow_b200._C import GlowkError
    conv = G.Conv2d(4, 8)
    x = torch.zeros(1, 4, 4, 4)
    with pytest.raises((GlowkError, NotImplementedError)):
        conv(x)                                       # CPU tensor: no fallback either way
    z = G.Conv2dZeros(4, 4)
    out = torch.nn.functional.conv2d(x, z.weight, z.bias, padding=1) * torch.exp(z.logs * 3)
    assert out.requires_grad                          # the torch formula the differentiable path uses


def test_flowmodel_training_under_dataparallel_replica_raises():
    import pytest
    import pytorch_glow_b200 as G
    fm = G.FlowModel((8, 8, 3), 8, K=1, L=1)
    fm._is_replica = True                             # what torch.nn.parallel.replicate sets on replicas
    with pytest.raises(RuntimeError, match="DataParallel"):
        fm.encode(torch.zeros(1, 3, 8, 8))


def test_layer_route_logdet_conventions_and_recompute_switch(monkeypatch):
    """Host logic of the stand-alone FlowStep route: the reference's logdet conventions (None / number / 0-dim / [N]
    tensor; an additive step keeps a scalar logdet scalar, an affine one returns one value per sample) and the
    activation-recompute switch read from the environment."""
    import importlib
    import pytorch_glow_b200 as G
    from pytorch_glow_b200 import config
    np.random.seed(0)
    add = G.FlowStep(4, 8, permutation="reverse", coupling="additive")
    aff = G.FlowStep(4, 8, permutation="reverse", coupling="affine")
    cpu = torch.device("cpu")
    vec, scalar_like = add._rows_logdet_in(0.5, 3, cpu)
    assert scalar_like and vec.shape == (3,) and torch.all(vec == 0.5)
    assert add._rows_logdet_out(vec + 1.0, scalar_like).dim() == 0            # additive: stays a scalar
    assert aff._rows_logdet_out(vec + 1.0, scalar_like).shape == (3,)         # affine: one value per sample
    vec, scalar_like = add._rows_logdet_in(torch.arange(3.), 3, cpu)
    assert not scalar_like and add._rows_logdet_out(vec, scalar_like).shape == (3,)
    assert add._rows_logdet_in(None, 3, cpu) == (None, False) and add._rows_logdet_out(None, False) is None
    assert not add._rows_route(torch.zeros(3, 4, 2, 2))                       # CPU tensors never take the rows route
    monkeypatch.setenv("GLOWK_RECOMPUTE", "1")
    assert importlib.reload(config).recompute_activations is True
    monkeypatch.setenv("GLOWK_RECOMPUTE", "0")
    assert importlib.reload(config).recompute_activations is False
