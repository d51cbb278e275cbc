"""reparameterised samples from the diagonal + rank-1 Gaussian q(v) (reference utils/sampler.py:4-21)"""
import torch


def sample_q_v(var_params_q_v, no_samples=1):
    mu, log_var, u = var_params_q_v['mu'], var_params_q_v['log_var'], var_params_q_v['u']
    sigma = torch.exp(0.5 * log_var)
    eps = torch.randn_like(sigma)
    x = torch.randn(1, device=u.device)
    delta = eps * sigma + x * u
    if no_samples == 1:
        return mu + delta
    return mu + delta, mu - delta  # antithetic pair
