"""
VIWarmStart -- the VI warm start of the reference (Trainer._run_VI, trainer/trainer.py:119-223) as a fused device path.

One `step()` == one VI iteration: two antithetic samples of q(v) = N(mu, diag(exp(log_var)) + u u^T) (utils/sampler.py:4-21)
go through the operators of the SGLD step as two chains on the shared mixture (Sobolev smoothing, integration, jittered warp,
residual map, virtual decimation, the mixture's Adam step per sample: trainer.py:79-117), the entropy terms and their
gradients are evaluated in closed form (model/loss.py:342-372), and the reference's Adam (optimizers/adam_rate_decay.py)
steps mu, log_var, u and the regulariser's hyper-parameters -- kernels of libirsgmcmc.so only (csrc/irs_vi.cu around
irs_sgld_step), every scalar on the device, capturable in a CUDA graph.  The drop-in path through autograd
(Trainer._run_VI(fused=False)) computes the same quantities with the module classes; tests/test_gpu_vi.py compares the two.
"""
import copy
import ctypes

import torch

from . import _lib
from .sampler import SGLDSampler


class VIWarmStart:
    def __init__(self, fixed, moving, var_params_q_v, config, device='cuda:0', lr_mu=0.01, lr_log_var=0.01, lr_u=0.01,
                 lr_decay=1e-3, betas=(0.9, 0.999), adam_eps=1e-8):
        cfg = copy.copy(config)
        cfg.tau = 0.0                    # no Langevin noise, no SGD update: the step only evaluates gradients
        cfg.hyper_mode = 'reference'     # the two samples step the shared mixture one after the other (trainer.py:135-136)
        self.sampler = s = SGLDSampler(fixed, moving, 2, cfg, device=device)
        self.device, self.lib = s.device, s.lib
        shape = (1,) + tuple(s.v.shape[1:])
        f32 = dict(device=s.device, dtype=torch.float32)
        self.mu, self.log_var, self.u = (var_params_q_v[k].detach().to(**f32).reshape(shape).contiguous().clone()
                                         for k in ('mu', 'log_var', 'u'))
        self._m = [torch.zeros(shape, **f32) for _ in range(3)]
        self._v = [torch.zeros(shape, **f32) for _ in range(3)]
        self._eps_store = torch.empty(shape, **f32)
        self.vi_state = torch.zeros(_lib.VI_STATE_SIZE, device=s.device, dtype=torch.float64)
        self._partials = torch.zeros(4 * 592, device=s.device, dtype=torch.float64)
        self._counter = torch.zeros(1, device=s.device, dtype=torch.int32)
        self.lr = (float(lr_mu), float(lr_log_var), float(lr_u))
        self.lr_decay, self.betas, self.adam_eps = float(lr_decay), betas, float(adam_eps)
        self._eps = self._x = None
        self._vib = None
        self._graph = None
        self.iteration = 0

    # -- explicit noise for parity tests -------------------------------------------------------------------------------
    def set_noise(self, eps=None, x=None, jitter_unit=None):
        """eps (1,3,...) N(0,1), x scalar N(0,1), jitter_unit (2,3,D,H,W) U[0,1) instead of the Philox streams"""
        self._eps = None if eps is None else eps.to(self.device, torch.float32).reshape(self.mu.shape).contiguous()
        self._x = None if x is None else torch.as_tensor(x, dtype=torch.float32).reshape(1).to(self.device)
        self.sampler.set_noise(None, jitter_unit)
        self._vib, self._graph = None, None

    def _buffers(self):
        b = _lib.ViBuffers()
        p = lambda t: None if t is None else t.data_ptr()
        b.mu, b.log_var, b.u = p(self.mu), p(self.log_var), p(self.u)
        for k in range(3):
            b.adam_m[k], b.adam_v[k] = p(self._m[k]), p(self._v[k])
        b.eps_store, b.vi_state, b.partials, b.counter = p(self._eps_store), p(self.vi_state), p(self._partials), p(self._counter)
        b.eps, b.x = p(self._eps), p(self._x)
        b.lr_mu, b.lr_log_var, b.lr_u = self.lr
        b.lr_decay, (b.beta1, b.beta2), b.adam_eps = self.lr_decay, self.betas, self.adam_eps
        return b

    def _enqueue(self):
        s = self.sampler
        if self._vib is None:
            self._vib = self._buffers()
        if s._cbuf is None:
            s._cbuf = s._buffers()
        _lib.check(self.lib.irs_vi_step(ctypes.byref(s._cconf), ctypes.byref(s._cbuf), ctypes.byref(self._vib), _lib.stream()))

    def step(self, n=1, use_graph=True):
        """n VI iterations; asynchronous"""
        done = 0
        if use_graph and n > 0:
            if self._graph is None:
                self._enqueue()              # warm-up outside capture
                torch.cuda.synchronize()
                done += 1
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue()
                self._graph = g
            while done < n:
                self._graph.replay()
                done += 1
        while done < n:
            self._enqueue()
            done += 1
        self.iteration += n
        self.sampler.iteration += n

    # -- checkpoint / resume (the reference has none: a VI run of 1024 iterations that dies starts over) -------------------------
    @torch.no_grad()
    def state_dict(self):
        """variational parameters, their Adam moments, the step counter / bias-correction products (`vi_state`) and the shared
        mixture / regulariser state with its Philox offset (`hyper`), as host tensors.  The noise of an iteration is a function of
        (seed, iteration counter on the device), so a warm start restored from it continues bit-identically."""
        return {'shape': tuple(self.mu.shape), 'seed': self.sampler.cfg.seed, 'iteration': self.iteration,
                'mu': self.mu.cpu(), 'log_var': self.log_var.cpu(), 'u': self.u.cpu(),
                'adam_m': [t.cpu() for t in self._m], 'adam_v': [t.cpu() for t in self._v],
                'vi_state': self.vi_state.cpu(), 'hyper': self.sampler.hyper.cpu()}

    @torch.no_grad()
    def load_state_dict(self, sd):
        """in place: device pointers (and with them a captured graph) stay valid"""
        if tuple(sd['shape']) != tuple(self.mu.shape) or sd['seed'] != self.sampler.cfg.seed:
            raise ValueError(f"checkpoint does not match this warm start: shape {tuple(sd['shape'])}, seed {sd['seed']} "
                             f"(here {tuple(self.mu.shape)}, {self.sampler.cfg.seed})")
        for name in ('mu', 'log_var', 'u', 'vi_state'):
            getattr(self, name).copy_(sd[name])
        for k in range(3):
            self._m[k].copy_(sd['adam_m'][k])
            self._v[k].copy_(sd['adam_v'][k])
        self.sampler.hyper.copy_(sd['hyper'])
        self.iteration = self.sampler.iteration = int(sd['iteration'])

    # -- results -------------------------------------------------------------------------------------------------------------
    def var_params(self):
        return {'mu': self.mu, 'log_var': self.log_var, 'u': self.u}

    def loss_terms(self):
        """the scalar terms of the last iteration (one device->host copy): per-sample data / regulariser terms as the SGLD
        step logs them, the two entropy terms, and the virtual decimation factors"""
        st, vs = self.sampler.stats.cpu(), self.vi_state.cpu()
        return {'data': st[:, _lib.STAT_DATA].float(), 'reg': st[:, _lib.STAT_REG].float(),
                'alpha': st[:, _lib.STAT_ALPHA].float(), 'reg_energy': st[:, _lib.STAT_ENERGY].float(),
                'entropy_sample': float(vs[_lib.VI_ENTROPY]), 'entropy_log_det': float(vs[_lib.VI_ENTROPY + 1]),
                'x': float(vs[_lib.VI_X])}
