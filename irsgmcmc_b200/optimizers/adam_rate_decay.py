"""
Adam with learning-rate decay lr / (1 + step * lr_decay), the optimiser of the mixture and regulariser
hyper-parameters (reference optimizers/adam_rate_decay.py:32-99; bias correction counted from the last re-init).
The fused sampler runs the same recurrence on the device (csrc/irs_hyper.cuh); this class serves the drop-in path.
"""
import math

import torch
from torch.optim import Optimizer


class Adam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, lr_decay=0.0, amsgrad=False):
        for name, val, ok in (('learning rate', lr, lr >= 0.0), ('lr_decay value', lr_decay, lr_decay >= 0.0),
                              ('epsilon value', eps, eps >= 0.0),
                              ('beta parameter at index 0', betas[0], 0.0 <= betas[0] < 1.0),
                              ('beta parameter at index 1', betas[1], 0.0 <= betas[1] < 1.0)):
            if not ok:
                raise ValueError('Invalid {}: {}'.format(name, val))
        super().__init__(params, dict(lr=lr, lr_decay=lr_decay, betas=betas, eps=eps, weight_decay=weight_decay,
                                      amsgrad=amsgrad))

    @torch.no_grad()
    def step(self, closure=None, reinit=False):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError('Adam does not support sparse gradients, please consider SparseAdam instead')
                grad = p.grad
                state = self.state[p]
                fresh = len(state) == 0
                if fresh:
                    state['step'] = 0
                if fresh or reinit:
                    state['reinit'] = state['step']
                    state['exp_avg'], state['exp_avg_sq'] = torch.zeros_like(p), torch.zeros_like(p)
                    if group['amsgrad']:
                        state['max_exp_avg_sq'] = torch.zeros_like(p)
                clr = group['lr'] / (1 + state['step'] * group['lr_decay'])
                state['step'] += 1
                t = state['step'] - state['reinit']
                bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
                if group['weight_decay'] != 0:
                    grad = grad.add(p, alpha=group['weight_decay'])
                m, v = state['exp_avg'], state['exp_avg_sq']
                m.mul_(b1).add_(grad, alpha=1 - b1)
                v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
                if group['amsgrad']:
                    torch.max(state['max_exp_avg_sq'], v, out=state['max_exp_avg_sq'])
                    v = state['max_exp_avg_sq']
                denom = (v.sqrt() / math.sqrt(bc2)).add_(group['eps'])
                p.addcdiv_(m, denom, value=-clr / bc1)
        return loss
