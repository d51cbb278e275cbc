"""
Data and regularisation losses with the reference's class names, constructor arguments, attributes and return values
(reference model/loss.py), computed by the CUDA kernels of libirsgmcmc.so:
  GMM.map        LCC normalisation: separable box filters over shared-memory tiles (+ adjoint boxes in backward)
  GMM.log_pdf    per-voxel K-component log-sum-exp with warp-shuffle reductions for the parameter gradients
  RegLoss        energy reduction of forward differences; backward = weighted 7-point stencil
SSD is the DataLoss the reference leaves to the user (SURVEY.md surprise 1): map = F - M with a single component.
"""
import math
from abc import ABC, abstractmethod

import numpy as np
import torch
from torch import nn
from torch.nn.functional import log_softmax

from . import distributions as model_distr
from .. import ops
from ..utils.diff_op import DifferentialOperator, GradientOperator


class DataLoss(nn.Module, ABC):
    """base class of data losses: residual map + reduction"""

    @abstractmethod
    def forward(self, z):
        pass

    @abstractmethod
    def map(self, im_fixed, im_moving):
        pass

    @abstractmethod
    def reduce(self, z):
        pass


class _LccNormalise(torch.autograd.Function):
    """(I - u) / sqrt(Box((I - u)^2) / k^3 + 1e-10),  u = Box(I) / k^3, replicate padding"""

    @staticmethod
    def forward(ctx, im, s):
        zn, a, rs = ops.lcc_normalise(im.contiguous(), s)
        ctx.save_for_backward(a, rs)
        ctx.s = s
        return zn

    @staticmethod
    def backward(ctx, g):
        a, rs = ctx.saved_tensors
        return ops.lcc_normalise_bwd(g.contiguous(), a, rs, ctx.s), None


class _MixtureLogPdf(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, log_std, logits):
        flat = z.reshape(-1).contiguous()
        logp, dz, _ = ops.gmm_log_pdf(flat, log_std, logits, want_dz=True)
        # snapshots: the shared mixture is stepped in place between a chain's forward pass and the final backward pass
        # (trainer/trainer.py:316-327 of the reference), and the gradient belongs to the values used in the forward pass
        ctx.save_for_backward(flat, dz, log_std.detach().clone(), logits.detach().clone())
        ctx.z_shape = z.shape
        return logp.view(1, -1)

    @staticmethod
    def backward(ctx, g):
        flat, dz, log_std, logits = ctx.saved_tensors
        g = g.reshape(-1).contiguous()
        K = log_std.numel()
        g_z = (g * dz).view(ctx.z_shape) if ctx.needs_input_grad[0] else None
        g_ls = g_lg = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            _, _, gp = ops.gmm_log_pdf(flat, log_std, logits, weights=g, want_param_grads=True)
            g_ls = gp[:K].to(log_std.dtype)
            # d logp / d logits_j = rho_j - pi_j, summed with weights g
            pi = torch.softmax(logits.detach() + 1e-2, dim=0)
            g_lg = (gp[8:8 + K] - pi.double() * g.double().sum()).to(logits.dtype)
        return g_z, g_ls, g_lg


class GMM(DataLoss):
    """Gaussian mixture negative log-likelihood of LCC-normalised residuals"""

    def __init__(self, no_components, s):
        super().__init__()
        self.no_components = no_components
        self.logits = nn.Parameter(torch.zeros(no_components))
        self.log_std = nn.Parameter(torch.zeros(no_components))
        self.register_buffer('_log_sqrt_2pi', torch.tensor(0.5 * math.log(2.0 * math.pi)))
        self.s = s
        self.kernel_sz = s * 2 + 1
        self.sz = float(self.kernel_sz ** 3)

    @torch.no_grad()
    def init_parameters(self, sigma):
        sigma = float(sigma)
        init = torch.linspace(math.log(sigma / 100.0), math.log(sigma * 5.0), steps=self.no_components)
        self.log_std.data.copy_(init)

    @property
    def log_proportions(self):
        return log_softmax(self.logits + 1e-2, dim=0)

    @property
    def log_scales(self):
        return self.log_std

    @property
    def proportions(self):
        return torch.exp(self.log_proportions)

    @property
    def scales(self):
        return torch.exp(self.log_scales)

    @property
    def precision(self):
        return torch.exp(-2.0 * self.log_std)

    def log_pdf(self, z):
        return _MixtureLogPdf.apply(z, self.log_std, self.logits)

    def log_pdf_VD(self, z_scaled):
        """The reference reaches this only from rescale_residuals (utils/util.py:330-347), whose inner autograd pass is
        replaced here by the closed form r = z^2 sum_k rho_k / sigma_k^2 evaluated in a kernel (ops.gmm_log_pdf).  There is
        no PyTorch path to fall back to."""
        raise NotImplementedError('GMM.log_pdf_VD: use utils.util.rescale_residuals (closed form, CUDA kernel)')

    def forward(self, z):
        return self.reduce(z)

    def map(self, im_fixed, im_moving):
        return _LccNormalise.apply(im_fixed, self.s) - _LccNormalise.apply(im_moving, self.s)

    def reduce(self, z):
        return -1.0 * self.log_pdf(z).sum()


class SSD(GMM):
    """sum of squared differences as a single zero-mean Gaussian with learnable scale: map = F - M, K = 1"""

    def __init__(self, no_components=1, s=0):
        super().__init__(1, 1)
        self.s = 0

    @torch.no_grad()
    def init_parameters(self, sigma):
        """the single component starts at the residuals' own scale.  (GMM.init_parameters with one component would return the
        first point of its linspace, sigma / 100: a precision of 1e4 / sigma^2 that throws a chain across the volume within
        a few transitions -- measured max |u_11| = 5 voxels after six transitions at 64^3.)"""
        self.log_std.data.fill_(math.log(float(sigma)))

    def map(self, im_fixed, im_moving):
        return im_fixed - im_moving


class _Energy(torch.autograd.Function):
    """y_c = sum |D v_c|^2 for the forward-difference operator; backward = 2 D^T D v as a stencil"""

    @staticmethod
    def forward(ctx, v):
        v_c = v.contiguous()
        ctx.save_for_backward(v_c)
        return ops.reg_energy(v_c).to(v.dtype)

    @staticmethod
    def backward(ctx, g_y):
        v, = ctx.saved_tensors
        return ops.reg_energy_grad(v, g_y.double().contiguous())


class RegLoss(nn.Module, ABC):
    """regularisation losses are functions of the energy y = sum |diff_op(v)|^2 per sample"""

    def __init__(self, diff_op=None, dims=None, learnable=False):
        super().__init__()
        self.dims = dims
        self.dof = np.prod(dims) * 3.0
        self.learnable = learnable
        if diff_op is None:
            self.diff_op = DifferentialOperator()
        elif isinstance(diff_op, str):
            self.diff_op = DifferentialOperator.from_string(diff_op)
        elif isinstance(diff_op, DifferentialOperator):
            self.diff_op = diff_op
        else:
            self.diff_op = diff_op()

    def forward(self, input, *args, **kwargs):
        if type(self.diff_op) is not GradientOperator:
            raise NotImplementedError('RegLoss: only GradientOperator has a CUDA kernel (the energy and its gradient are one '
                                      'fused stencil; the reference ships no other working operator)')
        y = _Energy.apply(input)   # fused: never materialises the (N,3,D,H,W,3) gradient tensor
        return self._loss(y, *args, **kwargs)

    @abstractmethod
    def _loss(self, y, *args, **kwargs):
        pass


class RegLoss_L2(RegLoss):
    """log-Gaussian prior: 0.5 w y - 0.5 dof log w"""

    def __init__(self, w_reg, diff_op=None, dims=None, learnable=False):
        super().__init__(diff_op=diff_op, dims=dims, learnable=learnable)
        self.log_w_reg = nn.Parameter(torch.tensor(math.log(w_reg)), requires_grad=learnable)

    def _loss(self, y):
        return 0.5 * self.log_w_reg.exp() * y - 0.5 * self.dof * self.log_w_reg, y.log()


class RegLoss_EnergyBased(RegLoss):
    """a prior on the scalar energy, converted to a prior on the field with dof degrees of freedom"""

    @abstractmethod
    def _mlog_energy_prior(self, y, *args, **kwargs):
        pass

    def _loss(self, y, *args, **kwargs):
        return self._mlog_energy_prior(y, *args, **kwargs) + (0.5 * self.dof - 1.0) * y.log(), y.log()


class RegLoss_LogNormal(RegLoss_EnergyBased):
    """log-normal prior on the energy; loc initialised at the mean of expGamma(dof / 2, w_reg / 2), scale = 4 loc"""

    def __init__(self, w_reg=1.0, diff_op=None, dims=None, learnable=False):
        super().__init__(diff_op=diff_op, dims=dims, learnable=learnable)
        loc_init = model_distr.LogEnergyExpGammaPrior(w_reg, self.dof).expectation().clone().detach()
        self.loc = nn.Parameter(loc_init, requires_grad=learnable)
        self.log_scale = nn.Parameter(math.log(4.0) + loc_init.log(), requires_grad=learnable)

    @property
    def scale(self):
        return self.log_scale.exp()

    def _mlog_energy_prior(self, y, *args, **kwargs):
        return y.log() + self.log_scale + 0.5 * ((y.log() - self.loc) / self.scale) ** 2


class Entropy(nn.Module, ABC):
    """base class for the entropy of a probability distribution"""

    @abstractmethod
    def forward(self, **kwargs):
        pass


class EntropyMultivariateNormal(Entropy):
    """
    entropy-related terms of the diagonal + rank-1 Gaussian q(v) = N(mu, diag(sigma^2) + u u^T) used by the VI warm start
    (reference model/loss.py:342-372): with (log_var, u) the log-determinant part; with (sample, mu, log_var, u) the
    Mahalanobis part via the Sherman-Morrison identity.  Elementwise torch expressions over the field (VI is section 8f
    "next": it runs on the same CUDA operators as the SGLD step through autograd).
    """

    def forward(self, **kwargs):
        if len(kwargs) == 2:
            log_var, u = kwargs['log_var'], kwargs['u']
            sigma = torch.exp(0.5 * log_var)
            dims = tuple(range(1, log_var.dim()))
            return 0.5 * (torch.log1p(torch.sum((u / sigma) ** 2, dim=dims)) + torch.sum(log_var, dim=dims))
        if len(kwargs) == 4:
            sample, mu, log_var, u = kwargs['sample'], kwargs['mu'], kwargs['log_var'], kwargs['u']
            sigma = torch.exp(0.5 * log_var)
            dims = tuple(range(1, log_var.dim()))
            sample_n, u_n = (sample - mu) / sigma, u / sigma
            t1 = torch.sum(sample_n ** 2, dim=dims)
            t2 = torch.sum(sample_n * u_n, dim=dims) ** 2 / (1.0 + torch.sum(u_n ** 2, dim=dims))
            return 0.5 * (t1 - t2)
        raise NotImplementedError
