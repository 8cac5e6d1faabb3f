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


def _gen_bfp_optim(optim, name):
    class BFPOptim(optim):
        """bfp_optim.py:10-58: the wrapped optimiser's fp32 update on the WIDE weights (kept in state['shadow_p']), then both BFP
        copies of the result: wide -> shadow_p, narrow -> p.data."""

        def __init__(self, *args, **kwargs):
            self.bfp_args = unpack_bfp_args(kwargs)
            self._bfp_modules = []
            super().__init__(*args, **kwargs)

        def attach_modules(self, modules):
            """Modules whose packed-weight caches must be dropped after every step (BFPLinear.invalidate_packed)."""
            self._bfp_modules = [m for m in modules if hasattr(m, "invalidate_packed")]
            return self

        def step(self, *args, **kwargs):
            if self.bfp_args['num_format'] == 'fp32':
                return super().step(*args, **kwargs)
            for group in self.param_groups:
                for p in group['params']:
                    if p.grad is None:
                        continue
                    state = self.state[p]
                    if 'shadow_p' not in state:
                        p.data.copy_(float_to_bfp_tiled(p.data, sgd_update=True, **self.bfp_args))
                    else:
                        p.data.copy_(state['shadow_p'])
            loss = super().step(*args, **kwargs)
            for group in self.param_groups:
                for p in group['params']:
                    if p.grad is None:
                        continue
                    state = self.state[p]
                    if 'shadow_p' not in state:
                        state['shadow_p'] = torch.zeros_like(p.data)
                    state['shadow_p'].copy_(float_to_bfp_tiled(p.data, sgd_update=True, **self.bfp_args))
                    p.data.copy_(float_to_bfp_tiled(p.data, **self.bfp_args))
            for m in self._bfp_modules:
                m.invalidate_packed()
            return loss

    BFPOptim.__name__ = "BFP" + name
    return BFPOptim


def get_bfp_optim(optim, name):
    """bfp_optim.py:60-64"""
    if name not in _bfp_optims:
        _bfp_optims[name] = _gen_bfp_optim(optim, name)
    return _bfp_optims[name]


class BFPAdam(torch.optim.Adam):
    """bfp_optim_lstm.py:12-95: Adam whose updated weight is constrained to the wide BFP format after every step.  The reference
    reads its BFP arguments from bfp_config.yaml (bfp_util.get_bfp_args); here they are passed in (`bfp_args`, the same dict)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, bfp_args=None):
        self.bfp_args = unpack_bfp_args(dict(bfp_args or {}))
        super().__init__(params, lr, betas, eps, weight_decay, amsgrad)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None:
                    continue
                grad = p.grad.data
                if grad.is_sparse:
                    raise RuntimeError('Adam does not support sparse gradients, please consider SparseAdam instead')
                amsgrad = group['amsgrad']
                state = self.state[p]
                if len(state) == 0:
                    state['step'] = 0
                    state['exp_avg'] = torch.zeros_like(p.data)
                    state['exp_avg_sq'] = torch.zeros_like(p.data)
                    if amsgrad:
                        state['max_exp_avg_sq'] = torch.zeros_like(p.data)
                exp_avg, exp_avg_sq = state['exp_avg'], state['exp_avg_sq']
                beta1, beta2 = group['betas']
                state['step'] += 1
                if group['weight_decay'] != 0:
                    grad = grad.add(p.data, alpha=group['weight_decay'])
                exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
                exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
                if amsgrad:
                    torch.max(state['max_exp_avg_sq'], exp_avg_sq, out=state['max_exp_avg_sq'])
                    denom = state['max_exp_avg_sq'].sqrt().add_(group['eps'])
                else:
                    denom = exp_avg_sq.sqrt().add_(group['eps'])
                step_size = group['lr'] * math.sqrt(1 - beta2 ** state['step']) / (1 - beta1 ** state['step'])
                if self.bfp_args['num_format'] == 'fp32':
                    p.data.addcdiv_(exp_avg, denom, value=-step_size)
                elif self.bfp_args['num_format'] == 'bfp':
                    p.data.copy_(float_to_bfp_blocked(p.data.addcdiv_(exp_avg, denom, value=-step_size), sgd_update=True, **self.bfp_args))
                else:
                    raise NotImplementedError('NumFormat not implemented')
        return loss
