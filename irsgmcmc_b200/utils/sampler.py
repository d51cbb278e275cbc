"""reparameterised samples from the diagonal + rank-1 Gaussian q(v) (reference utils/sampler.py:4-21)"""
import torch


def sample_q_v(var_params_q_v, no_samples=1):
    mu, log_var, u = var_params_q_v['mu'], var_params_q_v['log_var'], var_params_q_v['u']
    sigma = torch.exp(0.5 * log_var)
    eps = torch.randn_like(sigma)
    x = torch.randn(1, device=u.device)
    if no_samples == 1:
        return mu + eps * sigma + x * u          # the reference's operation order: (mu + eps sigma) + x u
    delta = eps * sigma + x * u
    return mu + delta, mu - delta  # antithetic pair


@torch.no_grad()
def draw_chain_states(MCMC_init, var_params_q_v, no_chains, chain_offset=0, no_chains_total=None, generator=None,
                      state_shape=None, device=None):
    """
    Trainer.__SGLD_init of the reference (trainer/trainer.py:585-611) for the chains [chain_offset, chain_offset + no_chains)
    of a run with no_chains_total chains: returns (v (no_chains, 3, ...), sigma (1, 3, ...) or None for ones).

    'VI': one sample_q_v draw per chain in chain order (randn_like(sigma), then randn(1)) -- a shard skips the numbers of
    the chains before it, so identically seeded ranks reproduce exactly the states a single process would give the same
    global chain ids.  'noise': one randn call for all chains in the reference; the shard is a slice of that tensor.
    `generator` None = the default generator of the device, like the reference.
    """
    total = chain_offset + no_chains if no_chains_total is None else int(no_chains_total)
    if MCMC_init == 'VI':
        mu, log_var, u = var_params_q_v['mu'], var_params_q_v['log_var'], var_params_q_v['u']
        dev = mu.device if device is None else device
        mu, log_var, u = (t.detach().to(dev, torch.float32) for t in (mu, log_var, u))
        sigma = torch.exp(0.5 * log_var)
        v = torch.empty((no_chains,) + tuple(sigma.shape[1:]), device=dev)
        for c in range(chain_offset + no_chains):
            eps = torch.randn(sigma.shape, device=dev, dtype=sigma.dtype, generator=generator)
            x = torch.randn(1, device=dev, generator=generator)
            if c >= chain_offset:
                v[c - chain_offset] = (mu + eps * sigma + x * u)[0]
        return v, sigma.contiguous()
    shape = tuple(state_shape if state_shape is not None else var_params_q_v['mu'].shape[1:])
    dev = device if device is not None else (var_params_q_v['mu'].device if var_params_q_v is not None else 'cpu')
    if MCMC_init == 'identity':
        return torch.zeros((no_chains,) + shape, device=dev), None
    if MCMC_init == 'noise':
        full = torch.randn((total,) + shape, device=dev, generator=generator)
        return full[chain_offset:chain_offset + no_chains].clone(), None
    raise ValueError(f'unknown MCMC_init: {MCMC_init}')
