"""Where does FlowStep round-trip error come from?  (GPU box)"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pytorch_glow_b200 as G
from pytorch_glow_b200 import functional as K
from parity_util import adopt, randomize_

dev = "cuda:0"
for mode in ("fp32", "bf16"):
    for inp in ("rand", "randn"):
        np.random.seed(1); torch.manual_seed(1)
        fs = G.FlowStep(12, 512, "invconv", "affine")
        adopt(fs, randomize_({k: v.clone() for k, v in fs.state_dict().items()}, 5, coupling_std=0.01))
        fs.conv_dtype = mode
        fs = fs.to(dev).eval()
        g = torch.Generator().manual_seed(2)
        x = (torch.rand(8, 12, 32, 32, generator=g) if inp == "rand" else torch.randn(8, 12, 32, 32, generator=g)).to(dev)
        with torch.no_grad():
            an = fs.actnorm
            b, l = an.bias.detach().reshape(-1), an.logs.detach().reshape(-1)
            w, winv, ld = fs.invconv.prepared(True)
            z = K.actnorm_mix(x, w, None, b, l, 3.0, False)
            xb = K.actnorm_mix(z, winv, None, b, l, 3.0, True)
            print(mode, inp, "mix roundtrip max abs err %.3e  |z|max %.3g  cond(W) %.3g  |W Winv - I| %.3e" % (
                float((xb - x).abs().max()), float(z.abs().max()), float(torch.linalg.cond(w.double())),
                float((w.double() @ winv.double() - torch.eye(12, device=dev, dtype=torch.float64)).abs().max())))
            sa, sb = {}, {}
            p3a = fs.f.tap_rows(z, mode, save=sa)
            p3b = fs.f.tap_rows(z, mode, save=sb)
            for kk in ("a1", "h1", "h2"):
                print("   %s identical: %s" % (kk, bool(torch.equal(sa[kk], sb[kk]))))
            zc = z.clone(); zc[:, 6:] = 123.0
            p3c = fs.f.tap_rows(zc, mode)
            print("   tap_rows independent of z2:", bool(torch.equal(p3a[:, :108], p3c[:, :108])))
            print("   tap_rows deterministic:", bool(torch.equal(p3a[:, :108], p3b[:, :108])), "|p3|max %.3g" % float(p3a[:, :108].abs().max()))
            c3 = fs.f[4]
            z2 = z.clone()
            K.coupling(p3a, c3.bias.detach(), c3.logs.detach().reshape(-1), z2, True, False)
            z3 = z2.clone()
            K.coupling(p3a, c3.bias.detach(), c3.logs.detach().reshape(-1), z3, True, True)
            print("   coupling roundtrip max abs err %.3e |z2'|max %.3g" % (float((z3 - z).abs().max()), float(z2.abs().max())))
            y, _ = fs(x, None)
            back, _ = fs(y.clone(), None, reverse=True)
            e = (back - x).abs()
            print("   full roundtrip max abs err %.3e at %s" % (float(e.max()), np.unravel_index(int(e.argmax()), e.shape)))
