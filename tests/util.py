import torch
import torch.nn.functional as F


def rel(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def smooth_field(shape, amp, seed=0, passes=3):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g)
    for _ in range(passes):
        x = F.avg_pool3d(F.pad(x, (1, 1, 1, 1, 1, 1), mode='replicate'), 3, 1)
    return x / x.abs().max() * amp


def three_numbers(new32, ref32, ref64):
    """SURVEY section 8c protocol: (e_new, e_ref, new-vs-ref32)"""
    return rel(new32, ref64), rel(ref32, ref64), rel(new32, ref32)
