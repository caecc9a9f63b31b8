import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pytorch_glow_b200 as G
from pytorch_glow_b200 import functional as K
from parity_util import adopt, randomize_
DEV = "cuda:0"
np.random.seed(1); torch.manual_seed(1)
fm = G.FlowModel((64, 64, 3), 512, K=4, L=3, permutation="invconv", coupling="affine")
sd = randomize_({k: v.clone() for k, v in fm.state_dict().items()}, 5, coupling_std=0.01)
adopt(fm, sd)
fm = fm.to(DEV).eval()
x = torch.rand(8, 3, 64, 64, generator=torch.Generator().manual_seed(2)).to(DEV)
with torch.no_grad():
    if len(sys.argv) > 1:
        z, ld = fm(x, torch.zeros(8, device=DEV))
    h = fm.layers[0](x)[0]
    for li, layer in enumerate(list(fm.layers)[1:5]):
        an = layer.actnorm
        b, l = an.bias.detach().reshape(-1), an.logs.detach().reshape(-1)
        w, winv, _ = layer.invconv.prepared(True)
        zf = K.actnorm_mix(h, w, None, b, l, 3.0, False)
        p3f = layer.f.tap_rows(zf, None)
        y, _ = layer(h, None)
        print("layer", li, "z1 preserved:", bool(torch.equal(y[:, :6], zf[:, :6])), "|y|max %.3g" % float(y.abs().max()),
              "|p3|max %.3g" % float(p3f[:, :108].abs().max()))
        yc = y.clone()
        p3r = layer.f.tap_rows(yc, None)
        print("   p3 fwd==rev:", bool(torch.equal(p3f[:, :108], p3r[:, :108])), "max diff %.3e" % float((p3f[:, :108] - p3r[:, :108]).abs().max()))
        c3 = layer.f[4]
        K.coupling(p3r, c3.bias.detach(), c3.logs.detach().reshape(-1), yc, True, True)
        print("   z2 recovered err %.3e" % float((yc - zf).abs().max()))
        xb = K.actnorm_mix(yc, winv, None, b, l, 3.0, True)
        e = (xb - h).abs()
        print("   x recovered err %.3e at %s ; |W Winv - I| %.3e cond %.3g" % (float(e.max()), np.unravel_index(int(e.argmax()), e.shape),
              float((w.double() @ winv.double() - torch.eye(12, device=DEV, dtype=torch.float64)).abs().max()), float(torch.linalg.cond(w.double()))))
        back, _ = layer(y.clone(), None, reverse=True)
        e = (back - h).abs()
        print("   layer() roundtrip err %.3e ; per-channel max:" % float(e.max()), [float("%.2g" % v) for v in e.amax(dim=(0, 2, 3)).tolist()])
        h = y
