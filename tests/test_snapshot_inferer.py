"""the reference's snapshot wire format (misc/util.py:289-376) and the batched inferer paths
(network/inferer.py).  The snapshot tests are host logic (CPU); the inferer tests need the GPU."""
import os

import numpy as np
import pytest
import torch

import pytorch_glow_b200 as G
from pytorch_glow_b200 import snapshot
from pytorch_glow_b200.hps import make_hps
from pytorch_glow_b200.inferer import Inferer
from parity_util import randomize_


def _glow(lu=False, batch=4, seed=0, classes=0):
    hps = make_hps((16, 16, 3), K=2, L=2, hidden_channels=16, coupling="affine", permutation="invconv", batch=batch)
    hps.ablation.lu_decomposition = lu
    if classes:
        hps.dataset.num_classes = classes
    np.random.seed(seed); torch.manual_seed(seed)
    g = G.Glow(hps)
    sd = g.state_dict()
    randomize_(sd, seed + 1)
    g.load_state_dict(sd)
    return hps, g


def test_snapshot_roundtrip_reference_format(tmp_path, golden_glow):
    hps, g = _glow()
    opt = torch.optim.Adam([p for n, p in g.named_parameters() if n != "h_top"], lr=1e-3)
    path = snapshot.save_model(str(tmp_path), 1234, g, opt, 5.5, is_best=True)
    assert os.path.basename(path) == "network-snapshot-001234.pth"
    assert os.path.exists(os.path.join(str(tmp_path), "network-snapshot-best.pth"))
    state = torch.load(path)
    assert set(state) == {"step", "graph", "optimizer", "criterion", "seconds"} and state["step"] == 1234
    # keys / shapes are the reference's own (golden fixture = a state_dict produced by the reference)
    ref = golden_glow.sd("invconv_affine/sd/")
    assert list(state["graph"].keys()) == list(ref.keys())
    assert all(tuple(state["graph"][k].shape) == tuple(ref[k].shape) for k in ref)
    _, g2 = _glow(seed=7)
    for m in g2.modules():
        if isinstance(m, G.ActNorm):
            assert not m.bias_inited
    st = snapshot.load_model(str(tmp_path), "latest", g2, opt)
    assert st["step"] == 1234 and st["seconds"] == 5.5
    for (k, a), (_, b) in zip(g.state_dict().items(), g2.state_dict().items()):
        assert torch.equal(a, b), k
    assert all(m.bias_inited and m.logs_inited for m in g2.modules() if isinstance(m, G.ActNorm))   # util.py:368
    snapshot.load_model(str(tmp_path), 1234, g2)
    snapshot.load_model(str(tmp_path), "best", g2)
    with pytest.raises(FileNotFoundError):
        snapshot.load_model(str(tmp_path), 99, g2)


def test_snapshot_lu_models_export_and_import_dense_weights(tmp_path):
    _, dense = _glow(lu=False, seed=3)
    _, lu = _glow(lu=True, seed=4)
    snapshot.save_model(str(tmp_path), 1, dense, None, 0.0, is_best=False)
    snapshot.load_model(str(tmp_path), 1, lu)                      # dense reference snapshot -> LU parameters
    sd = snapshot.reference_state_dict(lu)                         # ... and back to the reference's keys
    assert list(sd.keys()) == list(dense.state_dict().keys())
    for k, v in dense.state_dict().items():
        assert float((sd[k] - v).abs().max()) < 1e-4, k
    # a different per-device batch only changes the (all-zero) h_top
    _, other = _glow(batch=8, seed=5)
    snapshot.load_model(str(tmp_path), 1, other)
    assert other.h_top.shape[0] == 8 and float(other.h_top.abs().sum()) == 0.0


@pytest.mark.gpu
def test_inferer_batched_paths_match_per_sample_loops():
    dev = "cuda:0"
    hps, g = _glow(classes=3)
    g = g.to(dev).eval()
    g.set_actnorm_inited()
    g.flow.set_conv_dtype("fp32")
    inf = Inferer(hps, g, devices=[dev], data_device=dev)
    gen = torch.Generator().manual_seed(0)
    imgs = torch.rand(10, 3, 16, 16, generator=gen)
    torch.manual_seed(0)
    zs = inf.encode_batch(imgs)                                   # 10 images through batches of h_top.shape[0] = 4
    assert zs.shape == (10, 24, 4, 4)
    # the dequantisation noise makes encode stochastic at the 1/256 level: compare with a tolerance on z
    torch.manual_seed(0)
    z0 = inf.encode(imgs[0])
    assert float((z0 - zs[0]).abs().max()) < 0.2
    # attribute deltas: batched masked products == the reference's per-sample loop (without its len(batch) quirk)
    y = (torch.rand(8, 3, generator=gen) > 0.5).float()
    batches = [{"x": imgs[:4], "y_onehot": y[:4]}, {"x": imgs[4:8], "y_onehot": y[4:8]}]
    torch.manual_seed(1)
    delta = inf.compute_attribute_delta(batches)
    torch.manual_seed(1)
    pos = np.zeros((3, 24, 4, 4)); neg = np.zeros_like(pos); npos = np.zeros(3); nneg = np.zeros(3)
    with torch.no_grad():
        for b in batches:
            z, _, _ = g(b["x"].to(dev))
            z = z.cpu().numpy()
            for i in range(4):
                for c in range(3):
                    if b["y_onehot"][i, c] > 0:
                        pos[c] += z[i]; npos[c] += 1
                    else:
                        neg[c] += z[i]; nneg[c] += 1
    ref = np.stack([pos[c] / max(1.0, npos[c]) - neg[c] / max(1.0, nneg[c]) for c in range(3)])
    assert np.abs(delta - ref).max() < 1e-4
    torch.manual_seed(1)
    quirk = inf.compute_attribute_delta(batches, reference_quirk=True)    # first two samples of each batch only
    assert quirk.shape == delta.shape and np.abs(quirk - delta).max() > 0
    # interpolation sweep: one encode + batched decodes; shapes and finiteness (decode re-samples Split2d halves)
    alphas = np.stack([np.eye(3)[c] * a for c in range(3) for a in (-1.0, 0.0, 1.0)])
    out = inf.interpolate_batch(imgs[0], delta, alphas)
    assert out.shape == (9, 3, 16, 16) and bool(torch.isfinite(out).all())
    one = inf.apply_attribute_delta(imgs[0], delta, [0.5, 0.0, -0.5])
    assert one.shape == (3, 16, 16)
    assert inf.sample(eps_std=0.7).shape == (4, 3, 16, 16)


@pytest.mark.gpu
def test_inferer_matches_reference_inferer_fixture():
    """encode / decode / compute_attribute_delta (with the reference's len(batch) quirk) / apply_attribute_delta against
    outputs of the reference's own Inferer (network/inferer.py:62-188), generated by tests/golden/make_golden_inferer.py
    from the real reference on CPU.  L=1 model: decode is deterministic."""
    import os
    import pytorch_glow_b200 as G
    from pytorch_glow_b200.hps import make_hps
    from parity_util import adopt, GOLDEN
    z = np.load(os.path.join(GOLDEN, "inferer.npz"))
    dev = "cuda:0"
    ncls, B = 5, 4
    hps = make_hps((16, 16, 3), K=2, L=1, hidden_channels=16, coupling="affine", permutation="invconv", batch=B)
    hps.dataset.num_classes = ncls
    np.random.seed(0)
    glow = G.Glow(hps)
    adopt(glow, {k[3:]: torch.from_numpy(np.array(z[k])) for k in z.files if k.startswith("sd/")})
    glow = glow.to(dev).eval()
    glow.flow.set_conv_dtype("fp32")
    inf = Inferer(hps, glow, devices=[dev], data_device=dev)
    orig = torch.nn.init.uniform_
    torch.nn.init.uniform_ = lambda t, a=0., b=1.: t.fill_((a + b) / 2)      # the fixture's deterministic dequantisation
    try:
        img = torch.from_numpy(z["img"])
        zz = inf.encode(img)
        assert float((zz.cpu() - torch.from_numpy(z["z"])).abs().max()) < 1e-4 * float(np.abs(z["z"]).max())
        rec = inf.decode(torch.from_numpy(z["z"]))
        assert float((rec.cpu() - torch.from_numpy(z["rec"])).abs().max()) < 1e-4 * max(1.0, float(np.abs(z["rec"]).max()))
        batches = [{"x": torch.from_numpy(z["batch%d/x" % i]), "y_onehot": torch.from_numpy(z["batch%d/y_onehot" % i])}
                   for i in range(3)]
        delta = inf.compute_attribute_delta(batches, reference_quirk=True)
        assert np.abs(delta - z["deltaz"]).max() < 1e-4 * max(1.0, np.abs(z["deltaz"]).max())
        out = inf.apply_attribute_delta(img, z["deltaz"].astype(np.float32), z["alpha"])
        assert float((out.cpu() - torch.from_numpy(z["interp"])).abs().max()) < 2e-4 * max(1.0, float(np.abs(z["interp"]).max()))
        # the batched sweep gives the same image for the same interpolation vector
        sweep = inf.interpolate_batch(img, z["deltaz"].astype(np.float32), np.stack([z["alpha"], 0 * z["alpha"]]))
        assert float((sweep[0] - out).abs().max()) < 1e-5 * max(1.0, float(out.abs().max()))
    finally:
        torch.nn.init.uniform_ = orig
