"""Autograd bindings (hand-written backward kernels).  Filled in by the training milestone."""


def _todo(*a, **k):
    raise NotImplementedError("backward pass kernels are not wired yet; call under torch.no_grad()")


actnorm_autograd = invconv_autograd = split2d_autograd = flowstep_autograd = _todo


class _Todo:
    @staticmethod
    def apply(*a, **k):
        _todo()


PermuteFunction = SqueezeFunction = _Todo
