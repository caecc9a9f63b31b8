"""Batched inference paths on top of the flow engine (reference: network/inferer.py, infer.py).

Same method names and return conventions as the reference's `Inferer` (sample / encode / decode /
compute_attribute_delta / apply_attribute_delta), with the per-sample Python loops replaced by batched device
work: attribute means are masked matrix products on the GPU, and an interpolation sweep (infer.py:145-153 runs one
encode + one decode of a replicated batch PER OUTPUT IMAGE) is one encode plus reverse passes over batches of
distinct latents.  Nothing here is on the measured hot path; it only calls Glow.forward / reverse.
"""
import numpy as np
import torch


class Inferer:
    def __init__(self, hps, graph, devices=None, data_device=None):
        self.hps = hps
        self.graph = graph
        self.graph.eval()
        self.devices = devices
        self.data_device = data_device if data_device is not None else next(graph.parameters()).device
        self.batch_size = self.graph.h_top.shape[0]                       # inferer.py:35
        self.num_classes = getattr(getattr(hps, 'dataset', None), 'num_classes', 0)
        self.y_condition = hps.ablation.y_condition

    # -- inferer.py:40-60 (the image grid is left to the caller: torchvision is not a dependency here)
    def sample(self, z=None, y_onehot=None, eps_std=0.5):
        with torch.no_grad():
            return self.graph(z=z, y_onehot=y_onehot, eps_std=eps_std, reverse=True)

    def _as_batch(self, t):
        t = torch.as_tensor(t, dtype=torch.float32)
        if t.dim() == 3:
            t = t.unsqueeze(0)
        return t.to(self.data_device)

    # -- inferer.py:62-82: latent of ONE image (the reference replicates it batch_size times and keeps row 0)
    def encode(self, img):
        return self.encode_batch(self._as_batch(img))[0]

    def encode_batch(self, imgs):
        """Latents of a batch of distinct images [N,C,H,W] -> [N,C_top,H_top,W_top] (any N)."""
        with torch.no_grad():
            out = []
            for i in range(0, imgs.shape[0], self.batch_size):
                z, _, _ = self.graph(self._as_batch(imgs[i:i + self.batch_size]))
                out.append(z)
            return torch.cat(out, 0)

    # -- inferer.py:84-99
    def decode(self, z):
        return self.decode_batch(self._as_batch(z))[0]

    def decode_batch(self, zs, eps_std=None):
        """Images of a batch of distinct latents (any N); Split2d halves are re-sampled like in the reference."""
        with torch.no_grad():
            out = []
            for i in range(0, zs.shape[0], self.batch_size):
                out.append(self.graph(z=self._as_batch(zs[i:i + self.batch_size]).clone(), y_onehot=None,
                                      eps_std=eps_std, reverse=True))
            return torch.cat(out, 0)

    # -- inferer.py:101-152
    def compute_attribute_delta(self, batches, reference_quirk=False):
        """deltaz[cls] = mean z of samples with the attribute - mean z of samples without it.

        `batches` yields dicts with 'x' [B,C,H,W] and 'y_onehot' [B,num_classes].  reference_quirk=True reproduces
        inferer.py:136 (`for i in range(len(batch))` iterates over the dict's TWO keys, so only the first two
        samples of every batch are accumulated) for comparisons with numbers produced by the reference."""
        with torch.no_grad():
            shape = tuple(self.graph.flow.output_shapes[-1][1:])
            pos = torch.zeros(self.num_classes, int(np.prod(shape)), device=self.data_device, dtype=torch.float64)
            neg = torch.zeros_like(pos)
            n_pos = torch.zeros(self.num_classes, device=self.data_device, dtype=torch.float64)
            n_neg = torch.zeros_like(n_pos)
            for batch in batches:
                x = batch['x'].to(self.data_device)
                y = batch['y_onehot'].to(self.data_device)
                z, _, _ = self.graph(x)
                if reference_quirk:
                    z, y = z[:len(batch)], y[:len(batch)]
                zf = z.reshape(z.shape[0], -1).double()
                m = (y > 0).double()                                   # [B, classes]
                pos += m.t() @ zf
                neg += (1.0 - m).t() @ zf
                n_pos += m.sum(0)
                n_neg += (1.0 - m).sum(0)
            delta = pos / n_pos.clamp_min(1.0).unsqueeze(1) - neg / n_neg.clamp_min(1.0).unsqueeze(1)
            return delta.reshape(self.num_classes, *shape).float().cpu().numpy()

    # -- inferer.py:154-188
    def apply_attribute_delta(self, img, deltaz, interpolation):
        return self.interpolate_batch(img, deltaz, [interpolation])[0]

    def interpolate_batch(self, img, deltaz, interpolations, eps_std=None):
        """One encode, then batched decodes of z + sum_i deltaz[i]*alpha[i] for every interpolation vector alpha
        (infer.py:145-153 sweeps 40 attributes x 9 levels one image at a time)."""
        deltaz = torch.as_tensor(np.asarray(deltaz), dtype=torch.float32).to(self.data_device)
        alphas = torch.as_tensor(np.asarray(interpolations), dtype=torch.float32).to(self.data_device)
        assert alphas.dim() == 2 and alphas.shape[1] == deltaz.shape[0] == self.num_classes
        z = self.encode(img)
        zs = z.unsqueeze(0) + (alphas @ deltaz.reshape(deltaz.shape[0], -1)).reshape(-1, *z.shape)
        return self.decode_batch(zs, eps_std)
