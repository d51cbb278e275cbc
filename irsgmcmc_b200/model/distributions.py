"""
Scalar log-densities used as hyper-priors on the hot path, with the reference's class names and constructor arguments
(reference model/distributions.py).  They act on K = 4 mixture parameters or one or two regulariser scalars per
iteration: no volume work, plain torch (in the fused sampler their gradients are evaluated in closed form on the
device, csrc/irs_hyper.cuh).
"""
import math

import torch
from torch import nn


def _as_scalar_tensor(value, what):
    if torch.is_tensor(value):
        if value.numel() != 1:
            raise ValueError('Invalid tensor size. Expected 1, got: {}'.format(value.numel()))
        return value.clone().detach().reshape(())
    return torch.tensor(float(value)) if not isinstance(value, float) else torch.tensor(value)


class NormalDistribution(nn.Module):
    """x ~ N(loc, scale); forward = log pdf"""

    def __init__(self, loc=None, scale=None, learnable=False):
        super().__init__()
        loc = 0.0 if loc is None else loc
        scale = math.log(10) if scale is None else scale
        self.loc = nn.Parameter(_as_scalar_tensor(loc, 'loc'), requires_grad=learnable)
        self.log_scale = nn.Parameter(_as_scalar_tensor(scale, 'scale').log(), requires_grad=learnable)
        self.register_buffer('_log_sqrt_2pi', torch.tensor(0.5 * math.log(2.0 * math.pi)))

    def forward(self, x):
        return -0.5 * ((x - self.loc) * torch.exp(-self.log_scale)) ** 2 - self.log_scale - self._log_sqrt_2pi


def gamma_log_pdf(log_x, shape, rate):
    return shape * torch.log(rate) + (shape - 1) * log_x - rate * log_x.exp() - torch.lgamma(shape)


def expgamma_log_pdf(x, shape, rate):
    """log density of X = log Z, Z ~ Gamma(shape, rate)"""
    return gamma_log_pdf(x, shape, rate) + x


def expgamma_expectation(shape, rate):
    return torch.digamma(shape) - torch.log(rate)


class _GammaDistribution(nn.Module):
    """x ~ Gamma(shape, rate) evaluated at log x (building block)"""

    def __init__(self, shape=1e-3, rate=1e-3, shape_learnable=False, rate_learnable=False, learnable=False):
        super().__init__()
        self.shape = nn.Parameter(_as_scalar_tensor(shape, 'shape'), requires_grad=learnable and shape_learnable)
        self.rate = nn.Parameter(_as_scalar_tensor(rate, 'rate'), requires_grad=learnable and rate_learnable)

    def expectation(self):
        return self.shape / self.rate

    def forward(self, log_x):
        return gamma_log_pdf(log_x, self.shape, self.rate)


class ExpGammaDistribution(nn.Module):
    """distribution of X = log Z for Z ~ Gamma(shape, rate)"""

    def __init__(self, shape=1e-3, rate=1e-3, shape_learnable=False, rate_learnable=False, learnable=False):
        super().__init__()
        self.gamma_distribution = _GammaDistribution(shape, rate, shape_learnable, rate_learnable, learnable)

    def expectation(self):
        return expgamma_expectation(self.gamma_distribution.shape, self.gamma_distribution.rate)

    def forward(self, x):
        return self.gamma_distribution(x) + x


class DirichletPrior(nn.Module):
    """symmetric (or per-class) Dirichlet log density evaluated at log proportions"""

    def __init__(self, no_classes, alpha=None):
        super().__init__()
        alpha = 0.5 if alpha is None else alpha
        if torch.is_tensor(alpha):
            if len(alpha) != no_classes:
                raise ValueError('Invalid tensor size. Expected {}, got: {}'.format(no_classes, len(alpha)))
            conc = alpha.clone().detach().reshape(-1)
        else:
            conc = torch.full(size=[no_classes], fill_value=float(alpha))
        self.concentration = nn.Parameter(conc, requires_grad=False)

    def forward(self, log_proportions):
        c = self.concentration
        return (log_proportions * (c - 1.0)).sum(-1) + torch.lgamma(c.sum(-1)) - torch.lgamma(c).sum(-1)


class LogPrecisionExpGammaPrior(nn.Module):
    """Gamma hyper-prior on w_reg expressed on log w_reg"""

    def __init__(self, shape=1e-3, rate=1e-3, shape_learnable=False, rate_learnable=False, learnable=False):
        super().__init__()
        self.expgamma_distribution = ExpGammaDistribution(shape, rate, shape_learnable, rate_learnable, learnable)

    def forward(self, x):
        return self.expgamma_distribution(x)


class LogEnergyExpGammaPrior(nn.Module):
    """prior on the location (a log energy) of the log-normal energy prior: exp(loc) ~ Gamma(nu dof / 2, nu w_reg / 2)"""

    def __init__(self, w_reg, dof, nu=1.0, learnable=False):
        super().__init__()
        self.nu = nn.Parameter(torch.tensor(nu), requires_grad=learnable)
        self.register_buffer('w_reg', torch.tensor(w_reg))
        self.register_buffer('dof', torch.tensor(dof))

    def expectation(self):
        return expgamma_expectation(0.5 * self.nu * self.dof, 0.5 * self.nu * self.w_reg)

    def forward(self, log_energy):
        return expgamma_log_pdf(log_energy, 0.5 * self.nu * self.dof, 0.5 * self.nu * self.w_reg)


class LogScaleNormalPrior(nn.Module):
    """normal prior on a log scale parameter"""

    def __init__(self, loc, scale, learnable=False):
        super().__init__()
        self.normal = NormalDistribution(loc, scale, learnable)

    def forward(self, log_scale):
        return self.normal(log_scale)
