"""
Trainer -- the SGLD part of the reference's Trainer (reference trainer/trainer.py) on the fused CUDA step.

Kept from the reference: the method names and signatures on the hot path (`_SGLD_transition(fixed, moving, data_loss,
reg_loss) -> (loss_terms, output, aux)`, `_run_MCMC`, the chain / mixture initialisation), the attribute names read from
the `trainer` section of the JSON config (`no_chains`, `MCMC_init`, `no_iters_burn_in`, `no_samples_MCMC`,
`log_period_MCMC`, `uniform_noise`), the kept-sample rule and the folding guard.  Dropped (out of
scope): TensorBoard / NIfTI / VTK writers, MetricTracker, SimpleITK surface distances, the data loader.

What differs by design: one transition is one CUDA-graph replay with no host synchronisation; the per-iteration
`.item()` logging of the reference (trainer.py:391-412) is replaced by on-device statistics read back on request;
samples are folded into on-device Welford moments instead of a host buffer (trainer.py:365-366,428-430,458), and with
several ranks the chains are sharded and the moments merged with NCCL.
"""
import time

import numpy as np
import torch

from .. import parallel
from ..sampler import SGLDConfig, SGLDSampler
from ..utils.util import calc_DSC_GPU, calc_no_non_diffeomorphic_voxels


def sampler_config_from_json(config):
    """map the reference's JSON config (configs/*/config.json) onto SGLDConfig"""
    dl, reg = config['data_loss'], config['reg_loss']
    tr = config['trainer']
    sob = config.get('Sobolev_grad', {'enabled': False})
    opt_gmm = config.get('optimizer_GMM', {}).get('args', {})
    opt_reg = config.get('optimizer_reg', {}).get('args', {})
    kw = dict(
        data_loss='ssd' if dl['type'] == 'SSD' else 'lcc',
        no_components=dl['args'].get('no_components', 1), s=dl['args'].get('s', 2),
        reg_loss=reg['type'], w_reg=reg['args']['w_reg'], reg_learnable=reg['args'].get('learnable', False),
        sobolev_enabled=sob.get('enabled', False), sobolev_s=sob.get('s', 3), sobolev_lambda=sob.get('lambda', 0.5),
        tau=config['optimizer_SG_MCMC']['args']['lr'],
        uniform_noise=tr['uniform_noise']['enabled'], uniform_noise_magnitude=tr['uniform_noise'].get('magnitude', 0.1),
        virtual_decimation=config.get('virtual_decimation', True),
        lr_log_std=opt_gmm.get('lr_log_std', 0.2), lr_logits=opt_gmm.get('lr_logits', 0.2),
        lr_reg=opt_reg.get('lr_loc', opt_reg.get('lr_log_w_reg', 0.01)), lr_decay=opt_gmm.get('lr_decay', 1e-3))
    if 'data_loss_scale_prior' in config:
        a = config['data_loss_scale_prior']['args']
        kw['gmm_scale_prior'] = (a['loc'], a['scale'])
    if 'data_loss_proportion_prior' in config:
        kw['dirichlet_alpha'] = config['data_loss_proportion_prior']['args'].get('alpha', 0.5)
    if 'reg_loss_scale_prior' in config:
        a = config['reg_loss_scale_prior']['args']
        kw['reg_scale_prior'] = (a['loc'], a['scale'])
    if reg['type'] not in ('RegLoss_LogNormal', 'RegLoss_L2'):
        raise NotImplementedError(reg['type'])
    return SGLDConfig(**kw)


class Trainer:
    def __init__(self, config, fixed, moving, var_params_q_v=None, structures_dict=None, device=None,
                 chain_offset=None, logger=None):
        """
        config: the reference's JSON config as a dict;  fixed / moving: {'im','mask','seg'} with shape (1,1,D,H,W)
        With torch.distributed initialised, `no_chains` is the TOTAL number of chains; this rank owns a contiguous shard.
        """
        self.config = config
        tr = config['trainer']
        self.MCMC_init = tr['MCMC_init']
        self.no_chains_total = int(tr['no_chains'])
        self.no_samples_MCMC = int(tr['no_samples_MCMC'])
        self.no_iters_burn_in = int(tr['no_iters_burn_in'])
        self.log_period_MCMC = int(tr['log_period_MCMC'])
        self.add_noise_uniform = tr['uniform_noise']['enabled']
        self.alpha = tr['uniform_noise'].get('magnitude', 0.1)
        self.virutal_decimation = config.get('virtual_decimation', True)  # (sic) reference trainer/trainer.py:42
        self.Sobolev_grad = config.get('Sobolev_grad', {}).get('enabled', False)
        self.structures_dict = structures_dict or {}
        self.logger = logger

        offset, count = parallel.chain_shard(self.no_chains_total)
        if chain_offset is not None:
            offset = chain_offset
        self.no_chains = count
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self.device = device
        self.fixed, self.moving = fixed, moving
        self.var_params_q_v = var_params_q_v
        self.sampler = SGLDSampler(fixed, moving, self.no_chains, sampler_config_from_json(config), device=device,
                                   chain_offset=offset)
        self.dims = self.sampler.dims
        self.no_voxels = int(np.prod(self.dims))
        self.SGLD_params = None
        self.return_masked_residuals = False
        self.clone_outputs = False
        self._gmm_pushed = False

    # -- initialisation (reference trainer.py:529-547, 585-611) ----------------------------------------------------
    def _GMM_init(self, v_sample=None):
        if v_sample is None and self.var_params_q_v is not None:
            from ..utils.sampler import sample_q_v
            v_sample = sample_q_v({k: v.to(self.device) for k, v in self.var_params_q_v.items()})
        self.sampler.init_gmm(v_sample)
        self._gmm_pushed = True

    def _SGLD_init(self, var_params_q_v=None, generator=None):
        var_params_q_v = var_params_q_v or self.var_params_q_v
        self.sampler.init_chains(self.MCMC_init, var_params_q_v, generator=generator)
        self.SGLD_params = {'sigma': self.sampler.sigma, 'tau': self.sampler.cfg.tau}
        self.v_curr_state = self.sampler.v

    # -- parameter mirroring between the drop-in modules and the device-side hyper state -----------------------------
    def _push_hyper(self, data_loss, reg_loss):
        s, K = self.sampler, self.sampler.cfg.no_components
        if data_loss is not None:
            s.hyper[1:1 + K] = data_loss.log_std.detach().double().to(s.device)
            s.hyper[9:9 + K] = data_loss.logits.detach().double().to(s.device)
        if reg_loss is not None:
            if hasattr(reg_loss, 'loc'):
                s.hyper[50] = reg_loss.loc.detach().double()
                s.hyper[51] = reg_loss.log_scale.detach().double()
            else:
                s.hyper[50] = reg_loss.log_w_reg.detach().double()

    @torch.no_grad()
    def _pull_hyper(self, data_loss, reg_loss):
        s, K = self.sampler, self.sampler.cfg.no_components
        if data_loss is not None:
            data_loss.log_std.copy_(s.hyper[1:1 + K].to(data_loss.log_std.dtype))
            data_loss.logits.copy_(s.hyper[9:9 + K].to(data_loss.logits.dtype))
        if reg_loss is not None:
            if hasattr(reg_loss, 'loc'):
                reg_loss.loc.copy_(s.hyper[50].to(reg_loss.loc.dtype))
                reg_loss.log_scale.copy_(s.hyper[51].to(reg_loss.log_scale.dtype))
            else:
                reg_loss.log_w_reg.copy_(s.hyper[50].to(reg_loss.log_w_reg.dtype))

    # -- the hot path -------------------------------------------------------------------------------------------------
    def _SGLD_transition(self, fixed=None, moving=None, data_loss=None, reg_loss=None):
        """
        one SGLD transition of every chain on this GPU; same return structure as the reference (trainer.py:291-356):
          loss_terms {'data': [C], 'reg': [C]},
          output {'im_moving_warped', 'displacement', 'transformation', 'curr_state'},
          aux {'residuals', 'alpha': [C], 'reg_energy': [C]}
        Scalars are returned as device tensors (no host synchronisation).  `data_loss` / `reg_loss`, when given, are
        the drop-in GMM / RegLoss modules: their parameters are pushed to the device state before the first transition
        and refreshed after every transition.
        """
        s = self.sampler
        if not self._gmm_pushed and (data_loss is not None or reg_loss is not None):
            self._push_hyper(data_loss, reg_loss)
            self._gmm_pushed = True
        s.step(1)
        if data_loss is not None or reg_loss is not None:
            self._pull_hyper(data_loss, reg_loss)
        st = s.stats
        C = self.no_chains
        loss_terms = {'data': [st[c, 1].float() for c in range(C)], 'reg': [st[c, 2].float() for c in range(C)]}
        out = s.output()
        if self.clone_outputs:
            out = {k: v.detach().clone() for k, v in out.items()}
        residuals = s.z
        if self.return_masked_residuals:
            residuals = s.z[s.mask.bool().expand(C, -1, -1, -1, -1)].view(C, -1)
        aux = {'residuals': residuals, 'alpha': [st[c, 0].float() for c in range(C)],
               'reg_energy': [st[c, 3].float() for c in range(C)]}
        return loss_terms, out, aux

    # -- the loop around it (reference trainer.py:358-476) -------------------------------------------------------------
    def _run_MCMC(self, data_loss=None, reg_loss=None, speed_test_iters=100):
        """burn-in + sampling; returns {'mean','std_dev', 'im_mean','im_std', 'n', 'DSC', 'no_non_diffeomorphic_voxels',
        'samples_per_sec'}; posterior statistics are over the kept samples of ALL ranks"""
        if self.SGLD_params is None:
            self._SGLD_init()
        s = self.sampler
        if not self._gmm_pushed:
            if data_loss is not None or reg_loss is not None:
                self._push_hyper(data_loss, reg_loss)
                self._gmm_pushed = True
            else:
                self._GMM_init()
        dsc, folded = [], []
        total = self.no_iters_burn_in + self.no_samples_MCMC
        from ..utils.diff_op import GradientOperator
        diff_op = GradientOperator()
        it = 0
        while it < total:
            # run up to the next kept sample in one go (graph replays, no host involvement)
            nxt = it + 1
            while nxt <= total and not self._is_kept(nxt):
                nxt += 1
            nxt = min(nxt, total)
            s.step(nxt - it)
            it = nxt
            if self._is_kept(it):
                s.accumulate()
                T = s.transformation()
                if self.structures_dict and s.moving_seg is not None and 'seg' in self.fixed:
                    seg_w = s.warp_segmentation(transformation=T)
                    seg_f = self.fixed['seg'].to(s.device).expand(self.no_chains, -1, -1, -1, -1)
                    dsc.append(calc_DSC_GPU(self.no_chains, seg_f, seg_w, self.structures_dict))
                n_folded, _ = calc_no_non_diffeomorphic_voxels(T, diff_op)
                folded.append(n_folded)
                if (n_folded > 0.001 * self.no_voxels).any():  # reference trainer.py:441-445 (exits the process there)
                    raise RuntimeError(f'sample {it}: {n_folded} voxels where the sampled transformation is not '
                                       f'diffeomorphic')
        if data_loss is not None or reg_loss is not None:
            self._pull_hyper(data_loss, reg_loss)
        mom = s.posterior_moments()
        result = {'mean': mom['displacement_mean'], 'std_dev': mom['displacement_std'], 'im_mean': mom['im_mean'],
                  'im_std': mom['im_std'], 'n': mom['n'], 'DSC': dsc, 'no_non_diffeomorphic_voxels': folded}
        if speed_test_iters:  # the reference's built-in speed test: transitions + one segmentation warp each (:467-476)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(speed_test_iters):
                s.step(1)
                if s.moving_seg is not None:
                    s.warp_segmentation()
            torch.cuda.synchronize()
            result['samples_per_sec'] = self.no_chains * speed_test_iters / (time.perf_counter() - t0)
        return result

    def _is_kept(self, sample_no):
        """reference trainer.py:414-415"""
        return sample_no > self.no_iters_burn_in and (sample_no % self.log_period_MCMC == 0 or
                                                      sample_no == self.no_samples_MCMC)
