"""CPU oracle for the Glow flow hot path -- TEST INFRASTRUCTURE ONLY.

This file is a functional, fp32, CPU restatement of the arithmetic that
corenel/pytorch-glow performs on its flow path (network/module.py,
network/model.py, misc/ops.py).  It is what the CUDA path is checked against.
Nothing under ``pytorch_glow_b200/`` imports it; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference from /root/reference (in the build container) and records its outputs
on seeded inputs into ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks every function below against those fixtures and against the
known-answer vectors of SURVEY.md section 8(c).  One piece is unpinned by the
reference: the LU parameterisation (``lu_assemble``), which the reference
raises NotImplementedError for (network/module.py:336-337); it is pinned
indirectly by loading the assembled W into the reference's dense layer.

All citations are file:line in /root/reference.  Parameters are passed as a
flat ``dict[str, Tensor]`` that uses the reference's ``state_dict()`` key
names, so a reference snapshot can be fed in unchanged.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LOG_2PI = float(np.log(2 * np.pi))  # network/module.py:405


# --------------------------------------------------------------------------
# misc/ops.py
# --------------------------------------------------------------------------
def reduce_sum(t, dim, keepdim=False):
    """Sequential per-dimension sums in sorted dim order (misc/ops.py:40-73)."""
    dims = sorted([dim] if isinstance(dim, int) else list(dim))
    for d in dims:
        t = t.sum(dim=d, keepdim=True)
    if not keepdim:
        for cnt, d in enumerate(dims):
            t = t.squeeze(d - cnt)
    return t


def reduce_mean(t, dim, keepdim=False):
    """Sequential per-dimension means in sorted dim order (misc/ops.py:4-37)."""
    dims = sorted([dim] if isinstance(dim, int) else list(dim))
    for d in dims:
        t = t.mean(dim=d, keepdim=True)
    if not keepdim:
        for cnt, d in enumerate(dims):
            t = t.squeeze(d - cnt)
    return t


def split_channel(t, split_type="simple"):
    """misc/ops.py:95-113: 'simple' = halves, 'cross' = even/odd channels."""
    assert t.dim() == 4 and split_type in ("simple", "cross")
    nc = t.shape[1]
    if split_type == "simple":
        return t[:, : nc // 2], t[:, nc // 2:]
    return t[:, 0::2], t[:, 1::2]


def cat_channel(a, b):
    """misc/ops.py:116-127."""
    return torch.cat((a, b), dim=1)


def count_pixels(t):
    """misc/ops.py:130-140: H*W."""
    return int(t.shape[2] * t.shape[3])


# --------------------------------------------------------------------------
# ActNorm (network/module.py:9-149)
# --------------------------------------------------------------------------
def actnorm_init(x, scale=1.0, logscale_factor=3.0, batch_variance=False):
    """Data-dependent init (network/module.py:86-120).

    bias = -mean_{N,H,W}(x); logs = log(scale/(sqrt(mean((x+bias)^2))+1e-6))/factor.
    Returns (bias, logs) shaped [1,C,1,1].
    """
    bias = -1.0 * reduce_mean(x, [0, 2, 3], keepdim=True)
    xc = x + bias
    if batch_variance:
        var = torch.mean(xc ** 2).reshape(1, 1, 1, 1)  # module.py:112-113
    else:
        var = reduce_mean(xc ** 2, [0, 2, 3], keepdim=True)
    logs = torch.log(scale / (torch.sqrt(var) + 1e-6)) / logscale_factor
    return bias, logs


def actnorm_init_reverse(x, scale=1.0, logscale_factor=3.0, batch_variance=False):
    """Data-dependent init when the first training-mode call comes with reverse=True (network/module.py:143-146:
    actnorm_scale runs first and initialises logs from the RAW input (62-63, 106-120), then actnorm_center
    initialises the bias from the scaled tensor (44-45, 93-104)).  Returns (bias, logs) shaped [1,C,1,1]."""
    if batch_variance:
        var = torch.mean(x ** 2).reshape(1, 1, 1, 1)
    else:
        var = reduce_mean(x ** 2, [0, 2, 3], keepdim=True)
    logs = torch.log(scale / (torch.sqrt(var) + 1e-6)) / logscale_factor
    xs = x * torch.exp(-(logs * logscale_factor))
    bias = -1.0 * reduce_mean(xs, [0, 2, 3], keepdim=True)
    return bias, logs.expand(1, x.shape[1], 1, 1).clone()


def actnorm(x, bias, logs, logdet=None, reverse=False, logscale_factor=3.0):
    """network/module.py:34-84,122-149.  Out-of-place restatement.

    fwd: y=(x+bias)*exp(f*logs), logdet += HW*sum(f*logs);
    rev: x=y*exp(-f*logs)-bias, logdet -= HW*sum(f*logs).
    """
    assert x.dim() == 4 and x.shape[1] == bias.shape[1]
    ls = logs * logscale_factor
    if not reverse:
        y = (x + bias) * torch.exp(ls)
    else:
        y = x * torch.exp(-ls) - bias
    if logdet is not None:
        d = torch.sum(ls) * count_pixels(x)
        logdet = logdet - d if reverse else logdet + d
    return y, logdet


# --------------------------------------------------------------------------
# Invertible 1x1 conv and permutations (network/module.py:322-397)
# --------------------------------------------------------------------------
def invconv_init_weight(num_channels, rng=np.random):
    """network/module.py:339-342: Q of QR(randn) drawn from numpy's RNG."""
    w = np.linalg.qr(rng.randn(num_channels, num_channels))[0].astype("float32")
    return torch.from_numpy(w)


def invconv(x, weight, logdet=None, reverse=False):
    """network/module.py:344-369.  z[n,o,h,w] = sum_i W[o,i] x[n,i,h,w]."""
    dlogdet = torch.log(torch.abs(torch.det(weight))) * count_pixels(x)
    c = weight.shape[0]
    if not reverse:
        z = F.conv2d(x, weight.view(c, c, 1, 1))
        if logdet is not None:
            logdet = logdet + dlogdet
    else:
        z = F.conv2d(x, weight.inverse().view(c, c, 1, 1))
        if logdet is not None:
            logdet = logdet - dlogdet
    return z, logdet


def lu_assemble(p, l, u, sign_s, log_s):
    """W = P . L . (U + diag(sign_s*exp(log_s))), log|det W| = sum(log_s).

    The reference has no LU path (network/module.py:336-337 raises); this is
    the openai/glow parameterisation BASELINE.json's north_star names.  L is
    unit-lower (strict lower part of `l` is used), U strictly upper.
    """
    c = p.shape[0]
    l_mask = torch.tril(torch.ones(c, c, dtype=l.dtype), -1)
    lo = l * l_mask + torch.eye(c, dtype=l.dtype)
    up = u * l_mask.t() + torch.diag(sign_s * torch.exp(log_s))
    return p @ lo @ up, torch.sum(log_s)


def lu_factor(weight):
    """Dense W -> (P, L, U, sign_s, log_s) with W == P L (U+diag(s)) (import path)."""
    w = weight.double()
    plu, piv = torch.linalg.lu_factor(w)
    pm, lm, um = torch.lu_unpack(plu, piv)
    s = torch.diagonal(um)
    return (pm.float(), torch.tril(lm, -1).float(), torch.triu(um, 1).float(),
            torch.sign(s).float(), torch.log(torch.abs(s)).float())


def permutation_indices(num_channels, shuffle=False, rng=np.random):
    """network/module.py:385-390: reversed arange, optionally numpy-shuffled."""
    idx = np.arange(num_channels - 1, -1, -1, dtype=np.int64)
    if shuffle:
        rng.shuffle(idx)
    inv = np.zeros(num_channels, dtype=np.int64)
    for i in range(num_channels):
        inv[idx[i]] = i
    return idx, inv


def permute(x, indices, indices_inverse, reverse=False):
    """network/module.py:392-397: x[:, idx] (bit-exact gather)."""
    assert x.dim() == 4
    sel = indices_inverse if reverse else indices
    return x[:, torch.as_tensor(np.asarray(sel), dtype=torch.long)]


# --------------------------------------------------------------------------
# Squeeze2d (network/module.py:539-612)
# --------------------------------------------------------------------------
def squeeze2d(x, factor=2):
    """out[n, c*f*f + fh*f + fw, i, j] = x[n, c, i*f+fh, j*f+fw] (module.py:573-591)."""
    if factor == 1:
        return x
    n, c, h, w = x.shape
    assert h % factor == 0 and w % factor == 0
    x = x.reshape(n, c, h // factor, factor, w // factor, factor)
    x = x.permute(0, 1, 3, 5, 2, 4).contiguous()
    return x.reshape(n, c * factor * factor, h // factor, w // factor)


def unsqueeze2d(x, factor=2):
    """Inverse of squeeze2d (module.py:551-570)."""
    if factor == 1:
        return x
    n, c, h, w = x.shape
    f2 = factor * factor
    assert c >= f2 and c % f2 == 0
    x = x.reshape(n, c // f2, factor, factor, h, w)
    x = x.permute(0, 1, 4, 2, 5, 3).contiguous()
    return x.reshape(n, c // f2, h * factor, w * factor)


def squeeze2d_numpy(x, factor=2):
    """Index-formula restatement used to cross-check squeeze2d bit-exactly."""
    n, c, h, w = x.shape
    out = np.empty((n, c * factor * factor, h // factor, w // factor), x.dtype)
    for fh in range(factor):
        for fw in range(factor):
            out[:, fh * factor + fw::factor * factor] = x[:, :, fh::factor, fw::factor]
    return out


# --------------------------------------------------------------------------
# GaussianDiag (network/module.py:400-483)
# --------------------------------------------------------------------------
def gaussian_eps(shape_tensor, eps_std=None):
    """module.py:408-421.  NB `eps_std or 1.` maps 0/None to 1."""
    eps_std = eps_std or 1.0
    return torch.normal(mean=torch.zeros_like(shape_tensor),
                        std=torch.ones_like(shape_tensor) * eps_std)


def gaussian_logps(mean, logs, x):
    """module.py:437-451."""
    return -0.5 * (LOG_2PI + 2.0 * logs + ((x - mean) ** 2) / torch.exp(2.0 * logs))


def gaussian_logp(mean, logs, x):
    """module.py:453-467: per-sample sum over C,H,W (sequential dims)."""
    return reduce_sum(gaussian_logps(mean, logs, x), [1, 2, 3])


def gaussian_sample(mean, logs, eps_std=None, eps=None):
    """module.py:469-483.  `eps` (already scaled by eps_std) may be supplied."""
    if eps is None:
        eps = gaussian_eps(mean, eps_std)
    return mean + torch.exp(logs) * eps


# --------------------------------------------------------------------------
# Coupling network (network/module.py:188-319)
# --------------------------------------------------------------------------
def _same_pad(weight):
    return tuple((k - 1) // 2 for k in weight.shape[2:])  # module.py:209-212


def conv2d_actnorm(x, weight, an_bias, an_logs):
    """`Conv2d` with do_actnorm=True: conv(no bias, SAME) then ActNorm (module.py:243-260).

    The inner ActNorm is assumed initialised (eval semantics, module.py:93-94).
    """
    y = F.conv2d(x, weight, None, 1, _same_pad(weight))
    y, _ = actnorm(y, an_bias, an_logs)
    return y


def conv2d_zeros(x, weight, bias, logs, logscale_factor=3.0):
    """`Conv2dZeros`: conv(+bias, SAME zero pad) * exp(logs*factor) (module.py:286-297)."""
    y = F.conv2d(x, weight, bias, 1, _same_pad(weight))
    return y * torch.exp(logs * logscale_factor)


def f_net(x, p, prefix):
    """`f()` = Conv2d 3x3 + ReLU + Conv2d 1x1 + ReLU + Conv2dZeros 3x3 (module.py:300-319)."""
    h = conv2d_actnorm(x, p[prefix + "0.weight"], p[prefix + "0.actnorm.bias"], p[prefix + "0.actnorm.logs"])
    h = torch.relu(h)
    h = conv2d_actnorm(h, p[prefix + "2.weight"], p[prefix + "2.actnorm.bias"], p[prefix + "2.actnorm.logs"])
    h = torch.relu(h)
    return conv2d_zeros(h, p[prefix + "4.weight"], p[prefix + "4.bias"], p[prefix + "4.logs"])


# --------------------------------------------------------------------------
# FlowStep (network/model.py:10-173)
# --------------------------------------------------------------------------
def flowstep(x, logdet, p, prefix, permutation="invconv", coupling="additive",
             perm=None, reverse=False):
    """One step of flow, out-of-place (bit-identical to the in-place original).

    fwd (model.py:82-117): actnorm -> perm -> split -> coupling -> cat.
    rev (model.py:119-154): exact inverse in reverse order.
    `perm` = (indices, indices_inverse) for 'reverse'/'shuffle'.
    """
    assert x.shape[1] % 2 == 0  # model.py:169
    if not reverse:
        z, logdet = actnorm(x, p[prefix + "actnorm.bias"], p[prefix + "actnorm.logs"], logdet)
        if permutation == "invconv":
            z, logdet = invconv(z, p[prefix + "invconv.weight"], logdet, reverse=False)
        else:
            z = permute(z, perm[0], perm[1], reverse=False)
        z1, z2 = split_channel(z, "simple")
        if coupling == "additive":
            z2 = z2 + f_net(z1, p, prefix + "f.")
        else:
            h = f_net(z1, p, prefix + "f.")
            shift, scale = split_channel(h, "cross")
            scale = torch.sigmoid(scale + 2.0)
            z2 = (z2 + shift) * scale
            logdet = reduce_sum(torch.log(scale), [1, 2, 3]) + logdet
        return cat_channel(z1, z2), logdet
    z1, z2 = split_channel(x, "simple")
    if coupling == "additive":
        z2 = z2 - f_net(z1, p, prefix + "f.")
    else:
        h = f_net(z1, p, prefix + "f.")
        shift, scale = split_channel(h, "cross")
        scale = torch.sigmoid(scale + 2.0)
        z2 = z2 / scale - shift
        logdet = -reduce_sum(torch.log(scale), [1, 2, 3]) + logdet
    z = cat_channel(z1, z2)
    if permutation == "invconv":
        z, logdet = invconv(z, p[prefix + "invconv.weight"], logdet, reverse=True)
    else:
        z = permute(z, perm[0], perm[1], reverse=True)
    z, logdet = actnorm(z, p[prefix + "actnorm.bias"], p[prefix + "actnorm.logs"], logdet, reverse=True)
    return z, logdet


# --------------------------------------------------------------------------
# Split2d (network/module.py:486-536)
# --------------------------------------------------------------------------
def split2d_prior(z1, p, prefix):
    """module.py:497-509: (mean, logs) = cross-split of Conv2dZeros(z1)."""
    h = conv2d_zeros(z1, p[prefix + "conv2d_zeros.weight"], p[prefix + "conv2d_zeros.bias"],
                     p[prefix + "conv2d_zeros.logs"])
    return split_channel(h, "cross")


def split2d(x, logdet, p, prefix, reverse=False, eps_std=None, eps=None):
    """module.py:511-536.  fwd returns z1 only; rev re-samples z2 from the prior."""
    if not reverse:
        z1, z2 = split_channel(x, "simple")
        mean, logs = split2d_prior(z1, p, prefix)
        return z1, gaussian_logp(mean, logs, z2) + logdet
    mean, logs = split2d_prior(x, p, prefix)
    z2 = gaussian_sample(mean, logs, eps_std, eps)
    return cat_channel(x, z2), logdet


# --------------------------------------------------------------------------
# FlowModel (network/model.py:176-314)
# --------------------------------------------------------------------------
def flow_layout(in_shape, K, L):
    """Layer list and output shapes (model.py:236-261).

    Returns [(kind, C_in, H_in, W_in)] with kind in squeeze|step|split, and the
    output_shapes list ([-1, C, H, W] per layer).
    """
    nh, nw, nc = in_shape
    assert nc in (1, 3)
    layers, shapes = [], []
    for i in range(L):
        layers.append(("squeeze", nc, nh, nw))
        nc, nh, nw = nc * 4, nh // 2, nw // 2
        shapes.append([-1, nc, nh, nw])
        for _ in range(K):
            layers.append(("step", nc, nh, nw))
            shapes.append([-1, nc, nh, nw])
        if i < L - 1:
            layers.append(("split", nc, nh, nw))
            nc = nc // 2
            shapes.append([-1, nc, nh, nw])
    return layers, shapes


def flow_encode(z, logdet, p, in_shape, K, L, permutation="invconv", coupling="additive",
                perms=None, prefix="flow."):
    """`FlowModel.encode` (model.py:263-276).  perms: {layer_index: (idx, inv)}."""
    layers, _ = flow_layout(in_shape, K, L)
    for i, (kind, _, _, _) in enumerate(layers):
        lp = "%slayers.%d." % (prefix, i)
        if kind == "squeeze":
            z = squeeze2d(z, 2)
        elif kind == "step":
            z, logdet = flowstep(z, logdet, p, lp, permutation, coupling,
                                 None if perms is None else perms.get(i))
        else:
            z, logdet = split2d(z, logdet, p, lp)
    return z, logdet


def flow_decode(z, p, in_shape, K, L, permutation="invconv", coupling="additive",
                perms=None, eps_std=None, eps_list=None, prefix="flow."):
    """`FlowModel.decode` (model.py:278-294): logdet is reset to 0. per layer; returns x only.

    eps_list, if given, supplies the Split2d noise in decode order (deepest split first).
    """
    layers, _ = flow_layout(in_shape, K, L)
    k = 0
    for i in reversed(range(len(layers))):
        kind = layers[i][0]
        lp = "%slayers.%d." % (prefix, i)
        if kind == "squeeze":
            z = unsqueeze2d(z, 2)
        elif kind == "step":
            z, _ = flowstep(z, 0.0, p, lp, permutation, coupling,
                            None if perms is None else perms.get(i), reverse=True)
        else:
            e = None if eps_list is None else eps_list[k]
            k += 1
            z, _ = split2d(z, 0.0, p, lp, reverse=True, eps_std=eps_std, eps=e)
    return z


# --------------------------------------------------------------------------
# Glow wrapper (network/model.py:317-550) -- thin caller that defines bits/dim
# --------------------------------------------------------------------------
def glow_nll(x, noise, p, in_shape, K, L, permutation, coupling, n_bits_x=8, perms=None):
    """`Glow.normal_flow` (model.py:409-452) with the dequantisation noise supplied.

    z = x + noise (noise ~ U(0, 1/n_bins)); objective = -ln(n_bins)*D + logdet +
    logp_top(z); nll = -objective/(ln2*D) in bits/dim, per sample.  The top prior
    is N(0,1) (h_top is all-zero and learn_top/y_condition are off, model.py:362-379).
    """
    n_bins = 2 ** n_bits_x
    z = x + noise
    d = x.shape[1] * count_pixels(x)
    objective = torch.zeros_like(x[:, 0, 0, 0]) + float(-np.log(n_bins)) * d
    z, objective = flow_encode(z, objective, p, in_shape, K, L, permutation, coupling, perms)
    mean = torch.zeros_like(z)
    logs = torch.zeros_like(z)
    objective = objective + gaussian_logp(mean, logs, z)
    nll = (-objective) / float(np.log(2.0) * d)
    return z, nll


def glow_sample(z, p, in_shape, K, L, permutation, coupling, perms=None, eps_std=None,
                eps_list=None, top_shape=None):
    """`Glow.reverse_flow` (model.py:454-471).  z=None draws the top latent first."""
    if z is None:
        h = torch.zeros(top_shape)
        z = gaussian_sample(h, h, eps_std)
    return flow_decode(z, p, in_shape, K, L, permutation, coupling, perms, eps_std, eps_list)


def generative_loss(nll):
    """model.py:496-506."""
    return torch.mean(nll)


# --------------------------------------------------------------------------
# Trainer step arithmetic (network/trainer.py:138-150) -- used for the train oracle
# --------------------------------------------------------------------------
def noam_lr(base_lr, global_step, warmup_steps=4000, min_lr=None):
    """misc/lr_scheduler.py:18-37."""
    step_num = global_step + 1.0
    lr = base_lr * warmup_steps ** 0.5 * min(step_num * warmup_steps ** -1.5, step_num ** -0.5)
    if global_step >= warmup_steps and min_lr is not None:
        lr = max(lr, min_lr)
    return lr


def clip_grads_(grads, max_grad_clip=5.0, max_grad_norm=100.0):
    """clip_grad_value_ then clip_grad_norm_ (trainer.py:142-147). Returns pre-clip total norm."""
    if max_grad_clip is not None and max_grad_clip > 0:
        for g in grads:
            g.clamp_(-max_grad_clip, max_grad_clip)
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    if max_grad_norm is not None and max_grad_norm > 0:
        coef = torch.clamp(max_grad_norm / (total + 1e-6), max=1.0)
        for g in grads:
            g.mul_(coef)
    return total


def adam_step_(param, grad, m, v, step, lr, beta1=0.9, beta2=0.9999, eps=1e-8):
    """torch.optim.Adam (no weight decay, no amsgrad) as used by builder.py:10-13,91-93."""
    m.mul_(beta1).add_(grad, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    param.addcdiv_(m, denom, value=-lr / bc1)
