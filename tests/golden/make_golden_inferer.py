"""Golden fixture for the inference helpers: runs the REAL reference `Inferer` (network/inferer.py:40-188) on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_inferer.py

A tiny L=1 Glow (no Split2d, so decode is deterministic) with randomised zero-convs goes through the reference's
encode / decode / compute_attribute_delta / apply_attribute_delta; inputs, the state_dict and the outputs are written
to tests/golden/inferer.npz.  `DataLoader` is replaced by a stub that yields the recorded batches in order (the
reference shuffles; its quirk at inferer.py:136 -- only the first TWO samples of every batch are accumulated --
makes the result depend on the batch composition, so the batches are part of the fixture).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GLOW_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from network import inferer as rinf  # noqa: E402
from network import model as rmodel  # noqa: E402
from pytorch_glow_b200.hps import make_hps  # noqa: E402


def midpoint_uniform(t, a=0., b=1.):
    """Deterministic stand-in for the dequantisation noise (network/model.py:421-423 draws it with uniform_): the
    fixture and the test both use the interval's mid-point, so encode() is reproducible bit for bit."""
    return t.fill_((a + b) / 2)


def main():
    torch.set_num_threads(4)
    ncls, B = 5, 4
    hps = make_hps((16, 16, 3), K=2, L=1, hidden_channels=16, coupling="affine", permutation="invconv", batch=B,
                   devices=("cpu",))
    hps.dataset.num_classes = ncls
    hps.dataset.num_workers = 0
    np.random.seed(11); torch.manual_seed(11)
    graph = rmodel.Glow(hps)
    g = torch.Generator().manual_seed(12)
    for name, m in graph.named_modules():
        cls = m.__class__.__name__
        if cls == "Conv2dZeros":
            m.weight.data.normal_(0, 0.05, generator=g); m.bias.data.normal_(0, 0.05, generator=g); m.logs.data.normal_(0, 0.1, generator=g)
        elif cls == "ActNorm":
            m.bias.data.normal_(0, 0.2, generator=g); m.logs.data.normal_(0, 0.1, generator=g)
            m.bias_inited = True; m.logs_inited = True
    graph.eval()
    torch.nn.init.uniform_ = midpoint_uniform           # (after the constructors, which draw their weights with it)
    inf = rinf.Inferer(hps, graph, devices=["cpu"], data_device="cpu")
    out = {"sd/" + k: v.detach().numpy().copy() for k, v in graph.state_dict().items()}
    img = torch.rand(3, 16, 16, generator=g)
    z = inf.encode(img.clone())
    rec = inf.decode(z.clone())
    out.update(img=img.numpy(), z=z.numpy(), rec=rec.numpy())
    # attribute deltas over three recorded batches
    batches = [{"x": torch.rand(B, 3, 16, 16, generator=g), "y_onehot": (torch.rand(B, ncls, generator=g) > 0.5).float()}
               for _ in range(3)]
    for i, b in enumerate(batches):
        out["batch%d/x" % i] = b["x"].numpy().copy()
        out["batch%d/y_onehot" % i] = b["y_onehot"].numpy().copy()

    class FakeLoader:
        def __init__(self, dataset, **kw):
            pass

        def __iter__(self):
            return iter([{k: v.clone() for k, v in b.items()} for b in batches])

        def __len__(self):
            return len(batches)
    rinf.DataLoader = FakeLoader

    class NumpyZ:
        """inferer.py:139 does `ndarray += torch.Tensor`, which numpy >= 2 refuses (TypeError).  The harness hands the
        reference's loop its latents as numpy arrays; the loop itself runs unmodified."""

        def __init__(self, graph):
            self.graph, self.flow = graph, graph.flow

        def __call__(self, x):
            z, nll, y = self.graph(x)
            return z.numpy(), nll, y
    real_graph = inf.graph
    inf.graph = NumpyZ(real_graph)
    deltaz = inf.compute_attribute_delta(dataset=None)
    inf.graph = real_graph
    out["deltaz"] = np.asarray(deltaz, dtype=np.float64)
    alpha = np.array([0.5, -1.0, 0.0, 2.0, 0.25], dtype=np.float32)
    out["alpha"] = alpha
    out["interp"] = inf.apply_attribute_delta(img.clone(), deltaz.astype(np.float32), alpha).numpy()
    path = os.path.join(HERE, "inferer.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), len(out), "arrays")


if __name__ == "__main__":
    main()
