"""
The closed forms the fused VI iteration evaluates (csrc/irs_vi.cu, DESIGN section 4 "VI warm start as a device path") against autograd
of the oracle's restatement of the reference (model/loss.py:342-372, utils/sampler.py:4-21, trainer/trainer.py:130-157), in fp64 on
the CPU: the two antithetic samples share the sample term of the entropy, mu drops out of it, and the gradients of
    L = (f(mu + delta) + f(mu - delta)) / 2 - (e_sample(+) + e_sample(-)) / 2 - e_logdet,   delta = eps sigma + x u
with respect to (mu, log_var, u) are the expressions of vi_update_kernel for an arbitrary differentiable f.
"""
import torch

from oracle import sgld_oracle as O


def test_vi_closed_form_gradients_match_autograd():
    torch.manual_seed(0)
    shape = (1, 3, 5, 4, 6)
    dt = torch.float64
    mu = torch.randn(shape, dtype=dt, requires_grad=True)
    log_var = (torch.randn(shape, dtype=dt) * 0.3 - 1.0).requires_grad_(True)
    u = (torch.randn(shape, dtype=dt) * 0.2 + 0.1).requires_grad_(True)
    eps, x = torch.randn(shape, dtype=dt), torch.randn(1, dtype=dt)
    A = torch.randn(shape, dtype=dt)

    def f(v):     # stands for data + regulariser of one sample: any smooth function of the sample
        return (torch.sin(v) * A).sum() + 0.3 * (v ** 2).sum() + (v[..., 1:] * v[..., :-1]).sum()

    sigma = torch.exp(0.5 * log_var)
    delta = eps * sigma + x * u
    s1, s2 = mu + delta, mu - delta
    e1 = O.entropy_terms(log_var, u, sample=s1, mu=mu).sum()
    e2 = O.entropy_terms(log_var, u, sample=s2, mu=mu).sum()
    e0 = O.entropy_terms(log_var, u).sum()
    assert abs(float(e1 - e2)) < 1e-9 * abs(float(e1))                      # the antithetic samples share the sample term
    loss = 0.5 * (f(s1) + f(s2)) - 0.5 * (e1 + e2) - e0
    g_mu, g_lv, g_u = torch.autograd.grad(loss, (mu, log_var, u))

    # what the kernels compute: g_k = d f / d sample_k, the four sums, then the closed forms
    with torch.enable_grad():
        a1 = s1.detach().requires_grad_(True); a2 = s2.detach().requires_grad_(True)
        g1, = torch.autograd.grad(f(a1), a1); g2, = torch.autograd.grad(f(a2), a2)
    sg = sigma.detach(); un = (u / sigma).detach(); a = eps + x * un
    t1, s_su, s_uu, s_lv = (a * a).sum(), (a * un).sum(), (un * un).sum(), log_var.detach().sum()
    assert abs(float(0.5 * (t1 - s_su ** 2 / (1 + s_uu)) - e1)) < 1e-9 * abs(float(e1))
    assert abs(float(0.5 * (torch.log1p(s_uu) + s_lv) - e0)) < 1e-9 * abs(float(e0))
    inv = 1.0 / (1.0 + s_uu)
    q = a * x - s_su * (a + x * un) * inv + s_su ** 2 * un * inv ** 2 + un * inv
    half_diff = 0.5 * (g1 - g2)
    c_mu = 0.5 * (g1 + g2)
    c_u = x * half_diff - q / sg
    c_lv = 0.5 * eps * sg * half_diff + 0.5 * un * q - 0.5
    for name, got, want in (('mu', c_mu, g_mu), ('u', c_u, g_u), ('log_var', c_lv, g_lv)):
        err = float((got - want).norm() / want.norm())
        assert err < 1e-12, (name, err)
