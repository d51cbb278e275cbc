"""
irsgmcmc_b200 -- B200-native SGLD registration step of dgrzech/ir-sgmcmc (hand-written CUDA for sm_100a behind a C ABI,
PyTorch for device memory / streams / torch.distributed).  See DESIGN.md.
"""
__version__ = '0.1.0'

from . import torch_ops  # noqa: E402,F401  registers torch.ops.irsgmcmc.* (CUDA kernels only)
