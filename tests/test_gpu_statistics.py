"""
North-star correctness criterion 3: "posterior moments statistically consistent with the reference sampler's".

The CUDA sampler runs with its OWN noise (Philox4x32-10 + Box-Muller for the Langevin term, Philox uniforms for the
jitter) next to the oracle's restatement of the reference loop with torch noise (reference trainer/trainer.py:291-356,
kept-sample rule :414-430, statistics utils/util.py:114-120).  No noise is injected, so the two runs share nothing but
the initial state: agreement can only be statistical.

Test design (a paired test that does not need the chains to have mixed): both implementations start C chains from the
same C states and run the same schedule (burn-in, then one kept sample every `period` transitions).  Chain c of the CUDA
sampler and chain c of the oracle differ only by their noise, and the C pairs are independent replicates, so for any
per-voxel statistic s (time-average of the displacement, of the warped image, time-std of the displacement) the spread
of the differences d_c = s_cuda,c - s_oracle,c over the chains is its Monte-Carlo error -- with no assumption on the
autocorrelation inside a chain:
    t(voxel) = mean_c d_c / sqrt(var_c d_c / C)
Under "same sampler" t is Student-t with C - 1 degrees of freedom, E[t^2] = (C-1)/(C-3) = 1.4 for C = 8 (the oracle
against itself with two torch seeds measures 1.34 .. 1.38).  The stated Monte-Carlo bound: the mean of t^2 over all
voxels is below 2.0 and fewer than 0.5 % of the voxels exceed |t| > 6.  The protocol has power: a sampler whose noise
scale is 10 % off measures mean t^2 = 5 .. 8, 30 % off 20 .. 50 (oracle against a perturbed oracle; the CUDA sampler
with a perturbed scale is checked in test_statistical_test_has_power).  The pooled moments of
SGLDSampler.posterior_moments() (Welford on the device) must agree with the oracle's calc_posterior_statistics of its own
pooled samples within the same Monte-Carlo error.
"""
import math

import pytest
import torch

from oracle import sgld_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

N, C = 16, 8
BURN, KEPT, PERIOD = 60, 60, 4


def _setup(seed=7):
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    torch.manual_seed(seed)
    fixed, moving, vp = make_pair(N)
    sigma = torch.exp(0.5 * vp['log_var'])
    v0 = sigma * torch.randn(C, 3, N, N, N) + 0.1 * torch.randn(C, 1, 1, 1, 1)
    return fixed, moving, sigma, v0


def _chain_stats(disp_samples, im_samples):
    """per-chain statistics over the kept samples: samples are (kept, C, k, D, H, W)"""
    return {'disp_mean': disp_samples.mean(0), 'im_mean': im_samples.mean(0), 'disp_std': disp_samples.std(0)}


def run_cuda(fixed, moving, sigma, v0, tau_scale=1.0, seed=2024):
    from irsgmcmc_b200.sampler import SGLDSampler, SGLDConfig
    cfg = SGLDConfig(seed=seed)
    s = SGLDSampler(fixed, moving, C, cfg, device=DEV)
    s.set_state(v0, sigma)
    s.init_gmm(sigma_hat=0.7)
    if tau_scale != 1.0:   # a deliberately wrong sampler for the power check: Langevin noise of the wrong magnitude
        s.set_state(v0, sigma * tau_scale)
    s.step(BURN)
    disp, im = [], []
    for _ in range(KEPT):
        s.step(PERIOD)
        s.accumulate()
        disp.append(s.displacement.clone())
        im.append(s.im_warped.clone())
    torch.cuda.synchronize()
    return _chain_stats(torch.stack(disp).cpu().double(), torch.stack(im).cpu().double()), s.posterior_moments(), s


def run_oracle(fixed, moving, sigma, v0, seed=99):
    torch.manual_seed(seed)
    st = O.State(O.Config(), v0, sigma.expand(C, -1, -1, -1, -1), (N, N, N))
    st.init_gmm(0.7)
    for _ in range(BURN):
        O.sgld_transition(st, fixed, moving)
    disp, im = [], []
    for _ in range(KEPT):
        for _ in range(PERIOD):
            lt, out, aux, g = O.sgld_transition(st, fixed, moving)
        disp.append(out['displacement'].clone())   # the sample is the noisy smoothed state's displacement (trainer.py:303,418)
        im.append(out['im_moving_warped'].clone())
    disp, im = torch.stack(disp).double(), torch.stack(im).double()
    pooled = O.posterior_statistics(disp.reshape(-1, 3, N, N, N).float())
    pooled_im = O.posterior_statistics(im.reshape(-1, 1, N, N, N).float())
    return _chain_stats(disp, im), {'displacement_mean': pooled[0], 'displacement_std': pooled[1], 'im_mean': pooled_im[0],
                                    'im_std': pooled_im[1]}


def paired_se(a, b):
    d = a - b
    return torch.sqrt(d.var(0) / d.shape[0]).clamp_min(1e-12)


def t_stat(a, b):
    """paired t per voxel from per-chain statistics (C, ...): chain c of both runs starts from the same state"""
    return (a - b).mean(0) / paired_se(a, b)


@pytest.fixture(scope='module')
def runs(built):
    fixed, moving, sigma, v0 = _setup()
    cuda_stats, cuda_pooled, sampler = run_cuda(fixed, moving, sigma, v0)
    orc_stats, orc_pooled = run_oracle(fixed, moving, sigma, v0)
    return dict(fixed=fixed, moving=moving, sigma=sigma, v0=v0, cuda=cuda_stats, cuda_pooled=cuda_pooled,
                oracle=orc_stats, oracle_pooled=orc_pooled, n_kept=sampler.n_kept)


def test_posterior_moments_consistent_with_oracle_sampler(runs):
    assert runs['n_kept'] == C * KEPT
    mask = runs['fixed']['mask'][0, 0]
    for key in ('disp_mean', 'im_mean', 'disp_std'):
        t = t_stat(runs['cuda'][key], runs['oracle'][key])
        if key == 'im_mean':   # the warped image does not vary outside the head (background 0): no Monte-Carlo spread there
            t = t[:, mask]
        t2, tail = float((t ** 2).mean()), float((t.abs() > 6).double().mean())
        print(f'{key}: mean t^2 = {t2:.3f} (expected 1.4), frac |t| > 6 = {tail:.2e}, max |t| = {float(t.abs().max()):.2f}')
        assert t2 < 2.0 and tail < 5e-3, (key, t2, tail)


def test_pooled_welford_moments_within_monte_carlo_error(runs):
    """SGLDSampler.posterior_moments() (device Welford over every kept sample of every chain) against the oracle's
    calc_posterior_statistics of its own pooled samples: differences in units of the replicate standard error"""
    cp, op = runs['cuda_pooled'], runs['oracle_pooled']
    assert cp['n'] == C * KEPT
    for key, skey in (('displacement_mean', 'disp_mean'), ('im_mean', 'im_mean')):
        z = (cp[key].cpu().double() - op[key].double()) / paired_se(runs['cuda'][skey], runs['oracle'][skey])
        if key == 'im_mean':
            z = z[:, runs['fixed']['mask'][0, 0]]
        print(f'pooled {key}: mean z^2 =', float((z ** 2).mean()))
        assert float((z ** 2).mean()) < 2.0
    # the pooled std (within- plus between-chain variance): log-ratio against the replicate error of the per-chain time-stds
    ratio = (cp['displacement_std'].cpu().double() / op['displacement_std'].double()).log()
    rel_se = paired_se(runs['cuda']['disp_std'], runs['oracle']['disp_std']) / runs['oracle']['disp_std'].mean(0)
    print('pooled displacement std: mean log ratio =', float(ratio.mean()), 'mean |log ratio| =', float(ratio.abs().mean()),
          'replicate rel. se =', float(rel_se.mean()))
    assert abs(float(ratio.mean())) < 0.02 and float(ratio.abs().mean()) < 2.0 * float(rel_se.mean())


def test_statistical_test_has_power(runs):
    """the same protocol must REJECT a sampler whose Langevin noise is 30 % too large"""
    wrong, _, _ = run_cuda(runs['fixed'], runs['moving'], runs['sigma'], runs['v0'], tau_scale=1.3)
    t = t_stat(wrong['disp_std'], runs['oracle']['disp_std'])
    t2 = float((t ** 2).mean())
    print('wrong-noise sampler: mean t^2 of disp_std =', t2)
    assert t2 > 10.0


def test_philox_streams_differ_between_chains_and_seeds(runs):
    """chains are replicates only if their noise streams are distinct: chain-to-chain spread of the time-average is non-zero
    everywhere in the head, and a different seed gives a different trajectory"""
    spread = runs['cuda']['disp_mean'].std(0)
    assert float(spread.min()) > 0.0
    other, _, _ = run_cuda(runs['fixed'], runs['moving'], runs['sigma'], runs['v0'], seed=1)
    assert float((other['disp_mean'] - runs['cuda']['disp_mean']).abs().max()) > 1e-3
