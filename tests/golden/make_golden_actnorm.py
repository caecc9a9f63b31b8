"""Golden fixture for the ActNorm init variants no reference caller uses (batch_variance=True; first training-mode call
in the reverse direction): runs the REAL reference ActNorm (network/module.py:9-149) on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_actnorm.py      ->  tests/golden/actnorm_init.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GLOW_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REF)

from network import module as rmod  # noqa: E402


def main():
    g = torch.Generator().manual_seed(21)
    out = {}
    for c, hw in ((12, 8), (48, 4)):
        x = torch.randn(6, c, hw, hw, generator=g) * 1.7 + torch.randn(1, c, 1, 1, generator=g)
        for bv in (False, True):
            for rev in (False, True):
                if not bv and not rev:
                    continue                      # the default init is covered by layers.npz
                an = rmod.ActNorm(c, scale=1.3, logscale_factor=3., batch_variance=bv).train()
                ld = torch.zeros(6)
                y, ld = an(x.clone(), ld, reverse=rev)      # (clone: the reference scales in place)
                tag = "c%d_bv%d_rev%d/" % (c, bv, rev)
                out.update({tag + "x": x.numpy(), tag + "y": y.detach().numpy(), tag + "logdet": ld.detach().numpy(),
                            tag + "bias": an.bias.detach().numpy(), tag + "logs": an.logs.detach().numpy()})
    np.savez_compressed(os.path.join(HERE, "actnorm_init.npz"), **out)
    print("wrote actnorm_init.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
