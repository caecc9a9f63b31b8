"""Hyper-parameter container with attribute access (what the reference gets from `easydict`,
misc/util.py:18-29), plus the profile shapes of the BASELINE configs."""
import json


class Hps(dict):
    """dict with recursive attribute access; stays JSON-serialisable (util.py:184-185 dumps it)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, Hps):
            v = Hps(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def load_profile(path):
    with open(path) as fh:
        return Hps(json.load(fh))


def make_hps(image_shape=(64, 64, 3), K=32, L=3, hidden_channels=512, coupling="affine", permutation="invconv",
             batch=16, n_bits_x=8, lu_decomposition=False, devices=("cuda:0",), seed=2384):
    """The fields Glow reads (model.py:331-356,419,441), defaulting to profile/celeba.json:47-70."""
    return Hps({
        "model": {"image_shape": list(image_shape), "hidden_channels": hidden_channels, "K": K, "L": L,
                  "actnorm_scale": 1.0, "weight_y": 0.0, "n_bits_x": n_bits_x},
        "ablation": {"learn_top": False, "y_condition": False, "lu_decomposition": lu_decomposition,
                     "flow_permutation": permutation, "flow_coupling": coupling, "seed": seed,
                     "max_grad_clip": 5, "max_grad_norm": 100},
        "optim": {"num_batch_train": batch, "optimizer": "adam",
                  "optimizer_args": {"lr": 1e-3, "betas": [0.9, 0.9999], "eps": 1e-8, "weight_decay": 0},
                  "lr_scheduler": "noam", "lr_scheduler_args": {"warmup_steps": 4000, "min_lr": 1e-4}},
        "device": {"graph": list(devices), "data": [devices[0]]},
        "dataset": {"num_classes": 40},
    })
