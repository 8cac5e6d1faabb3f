"""Host-side mirror of the reference's optimiser wrappers ("sgd_update wide-mantissa weight update"):
`get_bfp_optim` (/root/reference/src/transformers/bfp/bfp_optim.py:8-64) and `BFPAdam` (bfp_optim_lstm.py:12-95).  Both are dead
code in the reference (nothing imports them; bfp_optim.py itself imports `float_to_bfp_tiled`, a name bfp_ops.py does not define --
this package provides it), so they are mirrored for completeness: the update runs in fp32 by the wrapped torch optimiser, the
weights are then constrained by the CUDA quantiser -- `sgd_update=True` selects `weight_mant_bits` (16-bit mantissas: m = 15) for
the copy the next update starts from, `mant_bits` for the copy the forward / backward passes see.

Weights are rewritten through `.data`, which does not bump the tensor version: `BFPLinear` never serves a cached packed weight while
the module trains (bfp_ops.BFPLinear._cacheable), and `invalidate_packed()` is called on every module handed to `attach_modules`.
"""
import math

import torch

from .bfp_ops import float_to_bfp_blocked, float_to_bfp_tiled, unpack_bfp_args

_bfp_optims = {}


def _trainable(optimizer):
    """(param, state) of every parameter that has a gradient this step."""
    for group in optimizer.param_groups:
        for p in group['params']:
            if p.grad is not None:
                yield p, optimizer.state[p]


def _gen_bfp_optim(optim, name):
    class BFPOptim(optim):
        """bfp_optim.py:10-58.  Two BFP copies of every weight: the WIDE one (state['shadow_p'], `weight_mant_bits`) is what the
        wrapped optimiser updates in fp32; the NARROW one (`mant_bits`) is what p.data holds between steps."""

        def __init__(self, *args, **kwargs):
            self.bfp_args = unpack_bfp_args(kwargs)
            self._bfp_modules = []
            super().__init__(*args, **kwargs)

        def attach_modules(self, modules):
            """Modules whose packed-weight caches must be dropped after every step (BFPLinear.invalidate_packed)."""
            self._bfp_modules = [m for m in modules if hasattr(m, "invalidate_packed")]
            return self

        def _wide(self, t):
            return float_to_bfp_tiled(t, sgd_update=True, **self.bfp_args)

        def step(self, *args, **kwargs):
            if self.bfp_args['num_format'] == 'fp32':
                return super().step(*args, **kwargs)
            # the update starts from the wide weights (first step: the constrained initial weights)
            for p, state in _trainable(self):
                p.data.copy_(state['shadow_p'] if 'shadow_p' in state else self._wide(p.data))
            loss = super().step(*args, **kwargs)
            # keep the wide result for the next step, hand the narrow one to the model
            for p, state in _trainable(self):
                if 'shadow_p' not in state:
                    state['shadow_p'] = torch.zeros_like(p.data)
                state['shadow_p'].copy_(self._wide(p.data))
                p.data.copy_(float_to_bfp_tiled(p.data, **self.bfp_args))
            for m in self._bfp_modules:
                m.invalidate_packed()
            return loss

    BFPOptim.__name__ = "BFP" + name
    return BFPOptim


def get_bfp_optim(optim, name):
    """bfp_optim.py:60-64: one wrapper class per name."""
    cls = _bfp_optims.get(name)
    if cls is None:
        cls = _bfp_optims[name] = _gen_bfp_optim(optim, name)
    return cls


class BFPAdam(torch.optim.Adam):
    """bfp_optim_lstm.py:12-95: Adam (the formulation of that file: eps added to sqrt(v) before the bias corrections are folded into
    the step size) whose updated weight is constrained to the wide BFP format after every step.  The reference reads its BFP arguments
    from bfp_config.yaml (bfp_util.get_bfp_args); here they are passed in (`bfp_args`, the same dict)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, bfp_args=None):
        self.bfp_args = unpack_bfp_args(dict(bfp_args or {}))
        super().__init__(params, lr, betas, eps, weight_decay, amsgrad)

    def _moments(self, p, group):
        state = self.state[p]
        if not state:
            state['step'] = 0
            for k in ('exp_avg', 'exp_avg_sq') + (('max_exp_avg_sq',) if group['amsgrad'] else ()):
                state[k] = torch.zeros_like(p.data)
        state['step'] += 1
        return state

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        fmt = self.bfp_args['num_format']
        if fmt not in ('fp32', 'bfp'):
            raise NotImplementedError('NumFormat not implemented')
        for group in self.param_groups:
            b1, b2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                g = p.grad.data
                if g.is_sparse:
                    raise RuntimeError('Adam does not support sparse gradients, please consider SparseAdam instead')
                state = self._moments(p, group)
                if group['weight_decay'] != 0:
                    g = g.add(p.data, alpha=group['weight_decay'])
                m, v = state['exp_avg'], state['exp_avg_sq']
                m.mul_(b1).add_(g, alpha=1 - b1)
                v.mul_(b2).addcmul_(g, g, value=1 - b2)
                if group['amsgrad']:
                    torch.max(state['max_exp_avg_sq'], v, out=state['max_exp_avg_sq'])
                    v = state['max_exp_avg_sq']
                denom = v.sqrt().add_(group['eps'])
                t = state['step']
                p.data.addcdiv_(m, denom, value=-group['lr'] * math.sqrt(1 - b2 ** t) / (1 - b1 ** t))
                if fmt == 'bfp':
                    p.data.copy_(float_to_bfp_blocked(p.data, sgd_update=True, **self.bfp_args))
        return loss
