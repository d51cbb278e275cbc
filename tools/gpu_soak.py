"""Soak run (development): thousands of graph replays of the transition in each hyper mode, the VI iteration, and the swapped image
pipeline; checks that everything stays finite, that the ticket / counters are left zero and that the displacement stays in the
small-displacement regime the bench is quoted on."""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from irsgmcmc_b200.sampler import SGLDSampler, SGLDConfig
from irsgmcmc_b200.vi import VIWarmStart
from irsgmcmc_b200.data_loader.synthetic import make_pair

dev = 'cuda:0'
n = 64
fixed, moving, vp = make_pair(n, device=dev)
for mode, data in (('reference', 'lcc'), ('per_chain', 'lcc'), ('frozen', 'lcc'), ('reference', 'ssd')):
    cfg = SGLDConfig(hyper_mode=mode, data_loss=data, reg_loss='RegLoss_LogNormal' if data == 'lcc' else 'RegLoss_L2',
                     w_reg=1.6 if data == 'lcc' else 1.4, reg_learnable=data == 'lcc')
    s = SGLDSampler(fixed, moving, 6, cfg, device=dev)
    s.init_chains('VI', vp, generator=torch.Generator(device=dev).manual_seed(1))
    s.init_gmm()
    s.capture(1); s.capture(25)
    t0 = time.time()
    s.step(3000)
    s.accumulate()
    torch.cuda.synchronize()
    dt = time.time() - t0
    ok = bool(torch.isfinite(s.v).all() and torch.isfinite(s.stats).all() and torch.isfinite(s.hyper).all())
    print(f'{mode:10s} {data}: 3000 transitions x 6 chains in {dt:.2f} s, finite={ok}, counters zero={int(s._counters.abs().sum()) == 0}, '
          f'max|u_11|={float(s._maxabs[11]):.3f}, max|v|={float(s.v.abs().max()):.2f}, alpha={s.stats[:, 0].tolist()[:2]}, '
          f'iteration counter={float(s.hyper.reshape(-1)[56]):.0f}')
    assert ok and int(s._counters.abs().sum()) == 0
w = VIWarmStart(fixed, moving, vp, SGLDConfig(), device=dev)
w.sampler.init_gmm()
w.step(1000)
torch.cuda.synchronize()
print('VI: 1000 iterations, finite =', bool(torch.isfinite(w.mu).all() and torch.isfinite(w.log_var).all() and torch.isfinite(w.u).all()),
      'max|mu| =', float(w.mu.abs().max()), 'mean log_var =', float(w.log_var.mean()), w.loss_terms())
# swapped image pipeline against direct loads, 200 steps with alternating pairs
pin = lambda x: x.contiguous().pin_memory()
pairs = [(pin(fixed['im'] * (1 - 0.05 * k)), pin(moving['im'] * (1 + 0.03 * k)), pin(fixed['mask'].view(torch.uint8))) for k in range(2)]
outs = []
for piped in (False, True):
    torch.manual_seed(0)
    s = SGLDSampler(fixed, moving, 2, SGLDConfig(), device=dev)
    s.set_state(0.5 * torch.randn(2, 3, n, n, n), torch.exp(0.5 * vp['log_var']))
    s.init_gmm(sigma_hat=0.7)
    if piped:
        s.prefetch_images(*pairs[0])
    for k in range(200):
        if piped:
            s.commit_images()
            s.prefetch_images(*pairs[(k + 1) % 2])
        else:
            s.load_images(*pairs[k % 2])
        s.step(1)
    torch.cuda.synchronize()
    outs.append((s.v.clone(), s.hyper.clone()))
print('swapped pipeline == direct loads after 200 steps:', torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]))
assert torch.equal(outs[0][0], outs[1][0])
print('soak OK')
