// irs_hyper.cuh -- the scalar part of the SGLD transition: virtual decimation factor, the per-chain Adam step on the
// mixture parameters and the Adam step on the regulariser's hyper-parameters.  A handful of doubles per chain, run by
// one thread on the device so that no scalar ever travels to the host (the reference syncs on .item() every iteration).
//
// reference: trainer/trainer.py:68-77,316-339,353-354; optimizers/adam_rate_decay.py:32-99;
//            utils/util.py:446-485; model/loss.py:61-69,197-198,264-312; model/distributions.py:56-58,111-112,209-211
#pragma once
#include "irs_common.cuh"
#include "../../include/irsgmcmc.h"

struct IrsHyperCfg {
    int K, virtual_decimation, reg_type, reg_learnable;
    double lr_log_std, lr_logits, lr_reg0, lr_reg1, lr_decay, beta1, beta2, eps;
    double gmm_prior_loc, gmm_prior_scale, dirichlet_alpha;
    double reg_prior_loc, reg_prior_scale, w_reg, dof, w_reg_prior_shape, w_reg_prior_rate;
    double n_mask;
    double reg_grad_scale;   // factor on the chain-summed hyper-gradients of the regulariser (1; 0.5 for the VI mean of two samples)
};

// layout of the per-chain reduction produced by the statistics pass
#define IRS_SUM_NLL 0      // -sum log pdf
#define IRS_SUM_RR 1       // sum r^2           (r = virtual-decimation residual, 0 off the mask)
#define IRS_SUM_RD 2       // sum r r(+1 along D)
#define IRS_SUM_RH 3
#define IRS_SUM_RW 4
#define IRS_SUM_RHO 5      // K slots: sum rho_k
#define IRS_SUM_Q (5 + IRS_MAX_K)   // K slots: sum rho_k z^2 prec_k
#define IRS_SUM_COUNT (5 + 2 * IRS_MAX_K)

IRS_HD double irs_round_f32(double x) { return (double)(float)x; }

// The mixture parameters, their Adam moments and the VD factor are fp32 tensors in the reference, and this code runs
// on ONE thread at the end of a reduction kernel: everything transcendental is done in fp32 (a chain of fp64 exp / log
// / pow calls costs tens of microseconds of pure latency); only the differences of the large fp64 sums stay in fp64.

// log pi = log_softmax(logits + 1e-2)   (reference model/loss.py:67-69)
IRS_HD void irs_log_proportions(const double* logits, int K, float* logpi) {
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) m = fmaxf(m, (float)logits[k] + 1e-2f);
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += expf((float)logits[k] + 1e-2f - m);
    const float lse = m + logf(s);
    for (int k = 0; k < K; ++k) logpi[k] = (float)logits[k] + 1e-2f - lse;
}

IRS_HD void irs_gmm_table(const double* log_std, const double* logits, int K, IrsGmm& g) {
    float logpi[IRS_MAX_K];
    irs_log_proportions(logits, K, logpi);
    g.K = K;
    for (int k = 0; k < IRS_MAX_K; ++k) {
        g.lw[k] = k < K ? logpi[k] - (float)log_std[k] : -INFINITY;
        g.prec[k] = k < K ? expf(-2.0f * (float)log_std[k]) : 0.f;
    }
}

// reference utils/util.py:446-485.  NaN when a lag-1 correlation is negative, like the reference: torch.clamp(max=1)
// propagates NaN, whereas fminf(1, NaN) returns 1 (IEEE minNum) -- hence the explicit select.  A correlation of exactly 0
// gives -log(0) = +inf -> clamped to 1, in both.
IRS_HD double irs_vd_alpha(const double* sums, double n_mask) {
    // corr_a = (sum r r_+ / n) / (sum r^2 / n): n cancels; one fp64 reciprocal instead of seven divisions (this runs on the
    // critical path between two chains)
    (void)n_mask;
    const double inv_rr = 1.0 / sums[IRS_SUM_RR];
    float prod = 1.f;
    for (int a = 0; a < 3; ++a) {
        const float corr = (float)(sums[IRS_SUM_RD + a] * inv_rr);
        const float t = -0.63661977236758134f * logf(corr);
        prod *= (t != t) ? t : fminf(1.f, t);
    }
    return (double)sqrtf(prod);
}

// one Adam update of a scalar parameter.  f32: the parameter and its moments are fp32 tensors in the reference
// (optimizers/adam_rate_decay.py:86-97: sqrt(v) / sqrt(bc2) + eps ;  p -= (clr / bc1) * m / denom)
IRS_HD void irs_adam_update(double& p, double& m, double& v, double g, double lr, double decay, double bc1, double bc2,
                            double b1, double b2, double eps, bool f32) {
    if (f32) {
        const float gf = (float)g;
        const float mf = (float)b1 * (float)m + (float)(1.0 - b1) * gf;
        const float vf = (float)b2 * (float)v + (float)(1.0 - b2) * gf * gf;
        const float denom = sqrtf(vf) / (float)sqrt(bc2) + (float)eps;
        const float step = (float)((lr / decay) / bc1);
        m = (double)mf; v = (double)vf;
        p = (double)((float)p - step * (mf / denom));
        return;
    }
    m = b1 * m + (1.0 - b1) * g;
    v = b2 * v + (1.0 - b2) * g * g;
    const double denom = sqrt(v) / sqrt(bc2) + eps;
    p -= (lr / decay) / bc1 * (m / denom);
}

// bias corrections 1 - beta^t without pow(): beta^t is kept as a running product next to the step counter
IRS_HD void irs_adam_bias(double* beta_pow, double b1, double b2, double step_before, double& bc1, double& bc2) {
    if (step_before == 0.0) { beta_pow[0] = 1.0; beta_pow[1] = 1.0; }
    beta_pow[0] *= b1; beta_pow[1] *= b2;
    bc1 = 1.0 - beta_pow[0]; bc2 = 1.0 - beta_pow[1];
}

// One Adam step on (log_std, logits) with loss  alpha * NLL - sum_k logN(log_std_k; loc, scale) - logDir(log pi; a)
// (reference trainer/trainer.py:68-77).  `sums` were reduced with the parameters currently in `hyper`.
// Split per parameter so that a warp can update the 2 K parameters side by side (one lane each): every quantity a lane needs
// is computed from the PRE-update values with the same operations in the same order as the serial loop below, so both give
// bit-identical results.
struct IrsAdamCtx {
    double decay, bc1, bc2, rho_total;
    float logpi[IRS_MAX_K];
};

// everything shared by the 2 K updates, from the pre-update state (does not write)
IRS_HD void irs_gmm_adam_context(const double* hyper, const IrsHyperCfg& cfg, const double* sums, IrsAdamCtx& ctx) {
    irs_log_proportions(hyper + IRS_HYPER_LOGITS, cfg.K, ctx.logpi);
    ctx.rho_total = 0.0;
    for (int k = 0; k < cfg.K; ++k) ctx.rho_total += sums[IRS_SUM_RHO + k];
    const double step0 = hyper[IRS_HYPER_GMM_STEP];
    ctx.decay = 1.0 + step0 * cfg.lr_decay;
    double b1p = hyper[IRS_HYPER_GMM_BETA_POW], b2p = hyper[IRS_HYPER_GMM_BETA_POW + 1];
    if (step0 == 0.0) { b1p = 1.0; b2p = 1.0; }
    ctx.bc1 = 1.0 - b1p * cfg.beta1;
    ctx.bc2 = 1.0 - b2p * cfg.beta2;
}

// update of parameter `idx`: 0 .. K-1 = log_std[k], K .. 2K-1 = logits[k]
IRS_HD void irs_gmm_adam_param(double* hyper, const IrsHyperCfg& cfg, const double* sums, double alpha, const IrsAdamCtx& ctx,
                               int idx) {
    const int K = cfg.K, k = idx < K ? idx : idx - K;
    if (idx < K) {
        double* ls = hyper + IRS_HYPER_LOG_STD;
        const double inv_s2 = 1.0 / (cfg.gmm_prior_scale * cfg.gmm_prior_scale);
        const double g = alpha * (sums[IRS_SUM_RHO + k] - sums[IRS_SUM_Q + k]) + (ls[k] - cfg.gmm_prior_loc) * inv_s2;
        irs_adam_update(ls[k], hyper[IRS_HYPER_M_LOG_STD + k], hyper[IRS_HYPER_V_LOG_STD + k], g, cfg.lr_log_std, ctx.decay,
                        ctx.bc1, ctx.bc2, cfg.beta1, cfg.beta2, cfg.eps, true);
    } else {
        double* lg = hyper + IRS_HYPER_LOGITS;
        const double pi_k = (double)expf(ctx.logpi[k]);
        const double g = alpha * (-sums[IRS_SUM_RHO + k] + pi_k * ctx.rho_total) - (cfg.dirichlet_alpha - 1.0) * (1.0 - K * pi_k);
        irs_adam_update(lg[k], hyper[IRS_HYPER_M_LOGITS + k], hyper[IRS_HYPER_V_LOGITS + k], g, cfg.lr_logits, ctx.decay,
                        ctx.bc1, ctx.bc2, cfg.beta1, cfg.beta2, cfg.eps, true);
    }
}

// step counter and the running powers of beta (after every parameter has been updated)
IRS_HD void irs_gmm_adam_advance(double* hyper, const IrsHyperCfg& cfg) {
    const double step0 = hyper[IRS_HYPER_GMM_STEP];
    double bc1, bc2;
    irs_adam_bias(hyper + IRS_HYPER_GMM_BETA_POW, cfg.beta1, cfg.beta2, step0, bc1, bc2);
    hyper[IRS_HYPER_GMM_STEP] = step0 + 1.0;
}

IRS_HD void irs_gmm_adam_step(double* hyper, const IrsHyperCfg& cfg, const double* sums, double alpha) {
    IrsAdamCtx ctx;
    irs_gmm_adam_context(hyper, cfg, sums, ctx);
    for (int idx = 0; idx < 2 * cfg.K; ++idx) irs_gmm_adam_param(hyper, cfg, sums, alpha, ctx, idx);
    irs_gmm_adam_advance(hyper, cfg);
}

// Regulariser: per-chain loss value, the coefficient c_c = dL/dy_c that multiplies dE/dv in the field gradient, and one
// Adam step on the hyper-parameters (reference model/loss.py:197-198,264-312; trainer/trainer.py:334-339,353-354).
// With the ExpGamma hyper-prior on log y the two dof/(2y) terms cancel exactly; the simplified
// form is evaluated so that no 3e6-sized fp terms are subtracted.
// per-chain part: loss value, field-gradient coefficient, contributions to the hyper-parameter gradients
IRS_HD void irs_reg_chain_terms(const double* hyper, const IrsHyperCfg& cfg, double* st, double& g0, double& g1) {
    const double y = st[IRS_STAT_ENERGY];
    if (cfg.reg_type == IRS_REG_LOGNORMAL) {
        const double loc = hyper[IRS_HYPER_REG_P], log_scale = hyper[IRS_HYPER_REG_P + 1];
        const double scale = exp(log_scale), s2 = scale * scale;
        const double ly = log(y), r = ly - loc;
        st[IRS_STAT_REG] = ly + log_scale + 0.5 * r * r / s2 + (0.5 * cfg.dof - 1.0) * ly;
        st[IRS_STAT_REG_COEF] = cfg.reg_learnable ? 0.5 * cfg.w_reg + r / (s2 * y) : (r / s2 + 0.5 * cfg.dof) / y;
        g0 = -r / s2;
        g1 = 1.0 - r * r / s2;
    } else {
        const double lw = hyper[IRS_HYPER_REG_P], w = exp(lw);
        st[IRS_STAT_REG] = 0.5 * w * y - 0.5 * cfg.dof * lw;
        st[IRS_STAT_REG_COEF] = 0.5 * w;
        g0 = 0.5 * w * y - 0.5 * cfg.dof;
        g1 = 0.0;
    }
}

// hyper-prior gradients + one Adam step, given the chain-summed gradients
IRS_HD void irs_reg_adam(double* hyper, const IrsHyperCfg& cfg, double g0, double g1) {
    if (!cfg.reg_learnable) return;
    const double step0 = hyper[IRS_HYPER_REG_STEP], decay = 1.0 + step0 * cfg.lr_decay;
    double bc1, bc2;
    irs_adam_bias(hyper + IRS_HYPER_REG_BETA_POW, cfg.beta1, cfg.beta2, step0, bc1, bc2);
    if (cfg.reg_type == IRS_REG_LOGNORMAL) {
        g1 += (hyper[IRS_HYPER_REG_P + 1] - cfg.reg_prior_loc) / (cfg.reg_prior_scale * cfg.reg_prior_scale);
        irs_adam_update(hyper[IRS_HYPER_REG_P], hyper[IRS_HYPER_REG_M], hyper[IRS_HYPER_REG_V], g0, cfg.lr_reg0, decay,
                        bc1, bc2, cfg.beta1, cfg.beta2, cfg.eps, false);
        irs_adam_update(hyper[IRS_HYPER_REG_P + 1], hyper[IRS_HYPER_REG_M + 1], hyper[IRS_HYPER_REG_V + 1], g1,
                        cfg.lr_reg1, decay, bc1, bc2, cfg.beta1, cfg.beta2, cfg.eps, false);
    } else {
        const double w = exp(hyper[IRS_HYPER_REG_P]);
        g0 += -(cfg.w_reg_prior_shape - cfg.w_reg_prior_rate * w);
        irs_adam_update(hyper[IRS_HYPER_REG_P], hyper[IRS_HYPER_REG_M], hyper[IRS_HYPER_REG_V], irs_round_f32(g0),
                        cfg.lr_reg0, decay, bc1, bc2, cfg.beta1, cfg.beta2, cfg.eps, true);
    }
    hyper[IRS_HYPER_REG_STEP] = step0 + 1.0;
}

// serial composition (host emulation; the device kernel spreads the chains over threads)
IRS_HD void irs_reg_hyper_step(double* hyper, const IrsHyperCfg& cfg, int C, double* stats) {
    double g0 = 0.0, g1 = 0.0;
    for (int c = 0; c < C; ++c) {
        double a, b;
        irs_reg_chain_terms(hyper, cfg, stats + (size_t)c * IRS_STAT_SIZE, a, b);
        g0 += a; g1 += b;
    }
    irs_reg_adam(hyper, cfg, g0, g1);
}
