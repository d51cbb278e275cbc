"""TEST INFRASTRUCTURE — not product code.

Imports the *unmodified* reference (dgrzech/ir-sgmcmc) from ``/root/reference`` (or ``$IRSGMCMC_REF``) so that
golden vectors can be generated from it and the restated oracle (``oracle/sgld_oracle.py``) can be pinned to it.

The reference only exists in the build container; it does not travel to the GPU box.  Nothing under
``irsgmcmc_b200/`` may import this module.  Only ``tests/`` and ``tests/golden/make_golden.py`` do.

The reference star-imports I/O and plotting packages that are absent here (SimpleITK, vtk, tvtk, nibabel, matplotlib,
seaborn, skimage: ``utils/util.py:6,13-14``, ``logger/*``).  None of them is on the SGLD hot path, so they are replaced by
empty stub modules before the import.
"""
import os
import sys
import types

REF_CANDIDATES = [os.environ.get('IRSGMCMC_REF', ''), '/root/reference']

_STUBS = ['SimpleITK', 'vtk', 'vtk.util', 'vtk.util.numpy_support', 'matplotlib', 'matplotlib.pyplot', 'mpl_toolkits',
          'mpl_toolkits.mplot3d', 'seaborn', 'nibabel', 'tvtk', 'tvtk.api', 'skimage', 'skimage.data']


def reference_path():
    for p in REF_CANDIDATES:
        if p and os.path.isfile(os.path.join(p, 'trainer', 'trainer.py')):
            return p
    return None


def available():
    return reference_path() is not None


_loaded = None


def load():
    """returns a namespace with the reference's modules; raises RuntimeError when the reference is absent"""
    global _loaded
    if _loaded is not None:
        return _loaded

    ref = reference_path()
    if ref is None:
        raise RuntimeError('reference not present (looked in $IRSGMCMC_REF and /root/reference)')

    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)

    sys.modules['vtk'].vtkStructuredPointsReader = object
    sys.modules['vtk.util.numpy_support'].vtk_to_numpy = lambda *a, **k: None
    sys.modules['mpl_toolkits.mplot3d'].Axes3D = object
    sys.modules['tvtk.api'].tvtk = object
    sys.modules['tvtk.api'].write_data = lambda *a, **k: None

    # the reference's top-level packages are called `utils`, `model`, `trainer`, ... -- make sure ours do not shadow
    clash = [m for m in ('utils', 'model', 'trainer', 'optimizers', 'base', 'logger', 'data_loader') if m in sys.modules]
    if clash:
        raise RuntimeError(f'modules {clash} already imported; cannot import the reference side by side')

    sys.path.insert(0, ref)
    try:
        import utils as r_utils
        import utils.util as r_util
        import utils.functions as r_functions
        import utils.registration as r_registration
        import utils.transformation as r_transformation
        import utils.diff_op as r_diff_op
        import utils.sampler as r_sampler
        import model.loss as r_loss
        import model.distributions as r_distr
        import optimizers as r_optim
        import trainer as r_trainer
    finally:
        sys.path.remove(ref)

    ns = types.SimpleNamespace(path=ref, utils=r_utils, util=r_util, functions=r_functions,
                               registration=r_registration, transformation=r_transformation, diff_op=r_diff_op,
                               sampler=r_sampler, loss=r_loss, distr=r_distr, optim=r_optim, trainer=r_trainer)
    _loaded = ns
    return ns


def make_trainer(ref, dims, no_chains, reg_type='RegLoss_LogNormal', w_reg=1.6, learnable=True, K=4, s=2,
                 sobolev_s=3, sobolev_lambda=0.5, tau=0.4, uniform_noise=None, virtual_decimation=True,
                 lr_gmm=0.2, lr_reg=0.01, lr_decay=1e-3, dtype=None, cps=None):
    """
    builds the reference Trainer without running __init__ (hard-coded cuda:0, TensorBoard, pandas MetricTracker:
    base/base_trainer.py:16,52-54) and wires it like parse_config.py:110-148,215-249 + trainer/trainer.py:21-42,568-583
    """
    import math
    import numpy as np
    import torch

    L, D = ref.loss, ref.distr
    t = ref.trainer.Trainer.__new__(ref.trainer.Trainer)
    t.device = 'cpu'
    t.no_chains = no_chains

    gmm = L.GMM(K, s)
    dof = float(np.prod(dims) * 3.0)
    losses = {'data': {'loss': gmm, 'scale_prior': D.LogScaleNormalPrior(0.0, 2.3),
                       'proportion_prior': D.DirichletPrior(K, 0.5)}, 'reg': {}}

    if reg_type == 'RegLoss_LogNormal':
        reg = L.RegLoss_LogNormal(w_reg=w_reg, diff_op='GradientOperator', dims=dims, learnable=learnable)
        if learnable:
            losses['reg']['loc_prior'] = D.LogEnergyExpGammaPrior(w_reg, dof)
            losses['reg']['scale_prior'] = D.LogScaleNormalPrior(2.8, 5.0)
    elif reg_type == 'RegLoss_L2':
        reg = L.RegLoss_L2(w_reg=w_reg, diff_op='GradientOperator', dims=dims, learnable=learnable)
        if learnable:
            shape = 0.5 * dof
            losses['reg']['w_reg_prior'] = D.LogPrecisionExpGammaPrior(shape=shape, rate=1.0 / shape)
    else:
        raise ValueError(reg_type)

    losses['reg']['loss'] = reg
    t.losses = losses
    t.diff_op = reg.diff_op
    # parse_config.py:100-108: the transformation module by name; SVFFD_3D with "cps" in configs/experiment5
    t.transformation_module = ref.transformation.SVF_3D(dims) if cps is None else ref.transformation.SVFFD_3D(dims, cps)
    t.registration_module = ref.registration.RegistrationModule()

    Adam = ref.optim.Adam
    t.optimizer_GMM = Adam([{'params': [gmm.log_std], 'lr': lr_gmm}, {'params': [gmm.logits], 'lr': lr_gmm}],
                           lr_decay=lr_decay)
    if learnable:
        if reg_type == 'RegLoss_LogNormal':
            t.optimizer_reg = Adam([{'params': [reg.loc], 'lr': lr_reg}, {'params': [reg.log_scale], 'lr': lr_reg}],
                                   lr_decay=lr_decay)
        else:
            t.optimizer_reg = Adam(reg.parameters(), lr=lr_reg, lr_decay=lr_decay)

    t.Sobolev_grad = True
    S, _ = ref.functions.Sobolev_kernel_1D(sobolev_s, sobolev_lambda)
    S = torch.from_numpy(S).float().unsqueeze(0)
    S = torch.stack((S, S, S), 0)
    t.padding = (sobolev_s,) * 6
    t.S = {'x': S.unsqueeze(2).unsqueeze(2), 'y': S.unsqueeze(2).unsqueeze(4), 'z': S.unsqueeze(3).unsqueeze(4)}

    t.add_noise_uniform = uniform_noise is not None
    if uniform_noise is not None:
        t.alpha = uniform_noise
    t.virutal_decimation = virtual_decimation  # (sic) trainer/trainer.py:42

    if dtype is not None:
        for m in (gmm, reg, t.transformation_module, *losses['data'].values(), *losses['reg'].values()):
            m.to(dtype)
        t.S = {k: v.to(dtype) for k, v in t.S.items()}

    t._tau = tau
    return t


def attach_state(t, v0, sigma, tau):
    """what Trainer.__SGLD_init does after drawing v (trainer/trainer.py:603-611)"""
    import torch
    t.v_curr_state = v0.clone().requires_grad_(True)
    t.SGLD_params = {'sigma': sigma, 'tau': tau}
    t.optimizer_SG_MCMC = torch.optim.SGD([t.v_curr_state], lr=tau)
