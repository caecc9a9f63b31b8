"""Snapshot wire format of pytorch-glow (misc/util.py:289-376) for the B200 engine.

`network-snapshot-%06d.pth` files written by the reference load into pytorch_glow_b200.Glow unchanged (same
state_dict keys and shapes; a dense `invconv.weight` is LU-factorised on load when the model was built with
`lu_decomposition=True`), and snapshots written here load back into the reference: the `graph` entry always
carries a dense `invconv.weight` unless `keep_lu=True`.
"""
import os
import shutil

import torch


def get_model_name(step):
    """misc/util.py: snapshot file name of a step."""
    return 'network-snapshot-{:06d}.pth'.format(step)


def get_best_model_name():
    return 'network-snapshot-best.pth'


def _unwrap(graph):
    return graph.module if hasattr(graph, "module") else graph      # DataParallel / DDP wrappers (util.py:312)


def reference_state_dict(graph, keep_lu=False):
    """state_dict() with the reference's keys: LU-parameterised 1x1 convs are exported as their dense weight."""
    g = _unwrap(graph)
    sd = {k: v.detach().clone() for k, v in g.state_dict().items()}
    if keep_lu:
        return sd
    from .module import Invertible1x1Conv
    dense = {}
    for name, m in g.named_modules():
        if isinstance(m, Invertible1x1Conv) and m.lu_decomposition:
            c = m.num_channels
            eye = torch.eye(c, dtype=torch.float64)
            lf = torch.tril(m.l.detach().double().cpu(), -1) + eye
            s = m.sign_s.detach().double().cpu() * torch.exp(m.log_s.detach().double().cpu())
            uf = torch.triu(m.u.detach().double().cpu(), 1) + torch.diag(s)
            dense[name] = (m.p.detach().double().cpu() @ lf @ uf).float()
    if not dense:
        return sd
    out = {}
    lu_keys = ('p', 'sign_s', 'l', 'u', 'log_s')
    for k, v in sd.items():                       # keep the reference's key order: `weight` where the LU keys were
        mod, _, leaf = k.rpartition('.')
        if mod in dense and leaf in lu_keys:
            if mod + '.weight' not in out:
                out[mod + '.weight'] = dense[mod]
            continue
        out[k] = v
    return out


def save_model(result_subdir, step, graph, optimizer, seconds, is_best, criterion_dict=None, keep_lu=False):
    """misc/util.py:289-327: {'step', 'graph', 'optimizer', 'criterion', 'seconds'} -> network-snapshot-%06d.pth
    (+ a copy as network-snapshot-best.pth)."""
    state = {
        'step': step,
        'graph': reference_state_dict(graph, keep_lu),
        'optimizer': optimizer.state_dict() if optimizer is not None else {},
        'criterion': {},
        'seconds': seconds,
    }
    if criterion_dict is not None:
        state['criterion'] = {k: v.state_dict() for k, v in criterion_dict.items()}
    save_path = os.path.join(result_subdir, get_model_name(step))
    torch.save(state, save_path)
    if is_best:
        shutil.copy(save_path, os.path.join(result_subdir, get_best_model_name()))
    return save_path


def load_model(result_subdir, step_or_model_path, graph, optimizer=None, criterion_dict=None, device=None):
    """misc/util.py:330-376.  `latest` is resolved to the highest-numbered snapshot of result_subdir (the
    reference sets model_path=None for it and crashes)."""
    model_path = step_or_model_path
    if isinstance(step_or_model_path, int):
        model_path = get_model_name(step_or_model_path)
    if step_or_model_path == 'best':
        model_path = get_best_model_name()
    if step_or_model_path == 'latest':
        names = sorted(n for n in os.listdir(result_subdir)
                       if n.startswith('network-snapshot-') and n[len('network-snapshot-'):-4].isdigit())
        if not names:
            raise FileNotFoundError('Failed to find model snapshot with latest')
        model_path = names[-1]
    if not os.path.exists(model_path):
        model_path = os.path.join(result_subdir, model_path)
        if not os.path.exists(model_path):
            raise FileNotFoundError('Failed to find model snapshot with {}'.format(step_or_model_path))
    if isinstance(device, int):
        device = 'cuda:{}'.format(device)
    state = torch.load(model_path, map_location=device)
    g = _unwrap(graph)
    sd = dict(state['graph'])
    own = g.state_dict()
    if 'h_top' in sd and 'h_top' in own and sd['h_top'].shape != own['h_top'].shape:
        sd['h_top'] = own['h_top']           # all-zero, sized by the per-device batch (model.py:350-356)
    g.load_state_dict(sd)
    g.set_actnorm_inited()
    from . import module as _module
    _module.bump_weight_generation()
    if optimizer is not None and state.get('optimizer'):
        optimizer.load_state_dict(state['optimizer'])
    if criterion_dict is not None:
        for k in criterion_dict.keys():
            criterion_dict[k].load_state_dict(state['criterion'][k])
    return state
