"""Prints bench-workload stage times and the per-step max |u| for a given size / chain count (development aid)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from irsgmcmc_b200.sampler import SGLDSampler, SGLDConfig
from irsgmcmc_b200.data_loader.synthetic import make_pair

ap = argparse.ArgumentParser()
ap.add_argument('--size', type=int, default=256)
ap.add_argument('--chains', type=int, default=1)
ap.add_argument('--steps', type=int, default=20)
a = ap.parse_args()
n, C = a.size, a.chains
fixed, moving, vp = make_pair(n)
s = SGLDSampler(fixed, moving, C, SGLDConfig(), device='cuda:0')
s.init_chains('VI', vp, generator=torch.Generator(device='cuda:0').manual_seed(123))
s.init_gmm()
s.step(5, use_graph=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); s.step(a.steps, use_graph=True); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(f'{n}^3 x {C} chains: {ms:.3f} ms/step, {C * n ** 3 / ms / 1e6:.3f} G voxel-steps/s')
print('maxabs per step', [round(float(x), 3) for x in s._maxabs[:12].tolist()])
st = s.profile_stages()
print({k: round(v, 4) for k, v in st.items()})
