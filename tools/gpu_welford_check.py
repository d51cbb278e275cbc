import torch, sys
sys.path.insert(0, '/root/repo')
from irsgmcmc_b200 import ops
dev='cuda:0'
torch.manual_seed(0)
for C, n in ((64, 128), (8, 128), (5, 21)):
    x = torch.randn(C, 3, n, n, n, device=dev) * 2 + 1
    mean, m2 = torch.zeros(3, n, n, n, device=dev), torch.zeros(3, n, n, n, device=dev)
    cnt = ops.welford_update(x, 0, mean, m2)
    cnt = ops.welford_update(x * 0.5, cnt, mean, m2)
    allx = torch.cat((x, x * 0.5), 0)
    print(C, n, float((mean - allx.mean(0)).abs().max()), float((ops.welford_std(m2, cnt) - allx.std(0)).abs().max()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): ops.welford_update(x, 0, mean, m2)
    e1.record(); torch.cuda.synchronize()
    print('  ms per update', e0.elapsed_time(e1) / 5, 'GB/s', x.numel() * 4 / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e9)
