// irs_vi.cu -- the VI warm start as a device path: antithetic sampling of q(v) = N(mu, diag(sigma^2) + u u^T), the entropy
// terms and their gradients in closed form, Adam on the field-sized variational parameters.
// (reference trainer/trainer.py:119-171, utils/sampler.py:4-21, model/loss.py:342-372, optimizers/adam_rate_decay.py:32-99)
//
// One iteration = vi_sample_kernel -> the fused SGLD-step operators on the two samples (irs_sgld_step_scaled: C = 2, tau = 0,
// the shared mixture stepped sample after sample like the reference) -> vi_update_kernel -> vi_advance_kernel.
//
// Closed forms.  With sigma = exp(log_var / 2), u_n = u / sigma, a = eps + x u_n the two samples are mu +- sigma a, so the
// sample term of the entropy (model/loss.py:360-372) is the same for both:
//     e1 = (t1 - s_su^2 / (1 + s_uu)) / 2,   t1 = sum a^2,  s_su = sum a u_n,  s_uu = sum u_n^2
// and the log-determinant term (:350-358) is e0 = (log1p(s_uu) + sum log_var) / 2.  With g_k = d(data_k + reg_k)/d sample_k:
//     dL/dmu      = (g_1 + g_2) / 2
//     dL/du       = x (g_1 - g_2) / 2 - q / sigma
//     dL/dlog_var = eps sigma (g_1 - g_2) / 4 + u_n q / 2 - 1 / 2
//     q = d(e1 + e0)/du_n = a x - s_su (a + x u_n) / (1 + s_uu) + s_su^2 u_n / (1 + s_uu)^2 + u_n / (1 + s_uu)
// (mu cancels in the entropy: sample - mu = +-sigma a.)  Checked against the reference's autograd through the drop-in modules in
// tests/test_gpu_vi.py.
#include "irs_kernels.cuh"

namespace {

#define IRS_VI_CHAIN 0xFFFFFFF0u   // Philox "chain" key of the variational noise (no SGLD chain uses it)

__device__ __forceinline__ float4 vi_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void vi_st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float vi_scalar_x(const irs_vi_buffers& vi, unsigned long long seed, unsigned long long iter) {
    if (vi.x != nullptr) return __ldg(vi.x);
    float e[3];
    irs_normal3(seed, 0xFFFFFFFFu, IRS_VI_CHAIN, iter, e);
    return e[0];
}

// samples v[0] = mu + delta, v[1] = mu - delta; keeps eps; reduces the four entropy sums (deterministic grid reduction)
__global__ void __launch_bounds__(256)
vi_sample_kernel(irs_vi_buffers vi, unsigned long long seed, float* __restrict__ v, long long Vs) {
    __shared__ double sh[4 * 32];
    __shared__ double total[4];
    const unsigned long long iter = (unsigned long long)vi.vi_state[IRS_VI_STEP];
    const float x = vi_scalar_x(vi, seed, iter);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const bool vec = (Vs % 4 == 0);
    const long long step = (long long)gridDim.x * blockDim.x * (vec ? 4 : 1);
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * (vec ? 4 : 1); i < Vs; i += step) {
        const int nv = vec ? 4 : 1;
        float e[4][3];
        if (vi.eps == nullptr) {
            for (int k = 0; k < nv; ++k) irs_normal3(seed, (uint32_t)(i + k), IRS_VI_CHAIN, iter, e[k]);
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const long long o = (long long)ch * Vs + i;
            float mu[4], lv[4], uu[4], ep[4], s1[4], s2[4];
            if (vec) {
                const float4 a = vi_ld4(vi.mu + o), b = vi_ld4(vi.log_var + o), c = vi_ld4(vi.u + o);
                mu[0] = a.x; mu[1] = a.y; mu[2] = a.z; mu[3] = a.w;
                lv[0] = b.x; lv[1] = b.y; lv[2] = b.z; lv[3] = b.w;
                uu[0] = c.x; uu[1] = c.y; uu[2] = c.z; uu[3] = c.w;
                if (vi.eps != nullptr) { const float4 q = vi_ld4(vi.eps + o); ep[0] = q.x; ep[1] = q.y; ep[2] = q.z; ep[3] = q.w; }
            } else {
                mu[0] = vi.mu[o]; lv[0] = vi.log_var[o]; uu[0] = vi.u[o];
                if (vi.eps != nullptr) ep[0] = vi.eps[o];
            }
            for (int k = 0; k < nv; ++k) {
                if (vi.eps == nullptr) ep[k] = e[k][ch];
                const float sigma = expf(0.5f * lv[k]);
                const float delta = ep[k] * sigma + x * uu[k];          // utils/sampler.py:17: eps sigma + x u
                s1[k] = mu[k] + delta;
                s2[k] = mu[k] - delta;
                const float un = uu[k] / sigma, a = ep[k] + x * un;
                acc[0] += a * a; acc[1] += a * un; acc[2] += un * un; acc[3] += lv[k];
            }
            if (vec) {
                vi_st4(v + o, make_float4(s1[0], s1[1], s1[2], s1[3]));
                vi_st4(v + 3 * Vs + o, make_float4(s2[0], s2[1], s2[2], s2[3]));
                vi_st4(vi.eps_store + o, make_float4(ep[0], ep[1], ep[2], ep[3]));
            } else {
                v[o] = s1[0]; v[3 * Vs + o] = s2[0]; vi.eps_store[o] = ep[0];
            }
        }
    }
    double blk[4];
    irs_block_sum<4>(acc, blk, sh);
    if (irs_grid_sum<4>(blk, vi.partials, vi.counter, total)) {
        if (threadIdx.x == 0) {
            double* st = vi.vi_state;
            st[IRS_VI_X] = (double)x;
            for (int k = 0; k < 4; ++k) st[IRS_VI_SUMS + k] = total[k];
            st[IRS_VI_ENTROPY] = 0.5 * (total[0] - total[1] * total[1] / (1.0 + total[2]));
            st[IRS_VI_ENTROPY + 1] = 0.5 * (log1p(total[2]) + total[3]);
        }
    }
}

// reference Adam on one element (optimizers/adam_rate_decay.py:86-97), fp32 like the parameter tensors
__device__ __forceinline__ float vi_adam(float p, float g, float& m, float& v, float b1, float b2, float inv_sqrt_bc2, float eps,
                                         float step) {
    m = m * b1 + (1.0f - b1) * g;
    v = v * b2 + (1.0f - b2) * g * g;
    const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
    return p - step * (m / denom);
}

__global__ void __launch_bounds__(256)
vi_update_kernel(irs_vi_buffers vi, const float* __restrict__ grad, long long Vs) {
    const double* st = vi.vi_state;
    const float x = (float)st[IRS_VI_X];
    const float s_su = (float)st[IRS_VI_SUMS + 1], inv1 = (float)(1.0 / (1.0 + st[IRS_VI_SUMS + 2]));
    const double step0 = st[IRS_VI_STEP];
    double b1p = st[IRS_VI_BETA_POW], b2p = st[IRS_VI_BETA_POW + 1];
    if (step0 == 0.0) { b1p = 1.0; b2p = 1.0; }
    const double bc1 = 1.0 - b1p * vi.beta1, bc2 = 1.0 - b2p * vi.beta2, decay = 1.0 + step0 * vi.lr_decay;
    const float b1 = (float)vi.beta1, b2 = (float)vi.beta2, eps = (float)vi.adam_eps;
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    const float st_mu = (float)((vi.lr_mu / decay) / bc1), st_lv = (float)((vi.lr_log_var / decay) / bc1),
                st_u = (float)((vi.lr_u / decay) / bc1);
    const long long N = 3 * Vs;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < N; o += (long long)gridDim.x * blockDim.x) {
        const float g1 = grad[o], g2 = grad[N + o];
        const float lv = vi.log_var[o], u = vi.u[o], ep = vi.eps_store[o];
        const float sigma = expf(0.5f * lv), un = u / sigma, a = ep + x * un;
        const float q = a * x - s_su * (a + x * un) * inv1 + s_su * s_su * un * inv1 * inv1 + un * inv1;
        const float half_diff = 0.5f * (g1 - g2);
        const float g_mu = 0.5f * (g1 + g2);
        const float g_u = x * half_diff - q / sigma;
        const float g_lv = 0.5f * ep * sigma * half_diff + 0.5f * un * q - 0.5f;
        float m, v;
        m = vi.adam_m[0][o]; v = vi.adam_v[0][o];
        vi.mu[o] = vi_adam(vi.mu[o], g_mu, m, v, b1, b2, inv_sqrt_bc2, eps, st_mu);
        vi.adam_m[0][o] = m; vi.adam_v[0][o] = v;
        m = vi.adam_m[1][o]; v = vi.adam_v[1][o];
        vi.log_var[o] = vi_adam(lv, g_lv, m, v, b1, b2, inv_sqrt_bc2, eps, st_lv);
        vi.adam_m[1][o] = m; vi.adam_v[1][o] = v;
        m = vi.adam_m[2][o]; v = vi.adam_v[2][o];
        vi.u[o] = vi_adam(u, g_u, m, v, b1, b2, inv_sqrt_bc2, eps, st_u);
        vi.adam_m[2][o] = m; vi.adam_v[2][o] = v;
    }
}

__global__ void vi_advance_kernel(double* st, double beta1, double beta2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double step0 = st[IRS_VI_STEP];
    double bc1, bc2;
    irs_adam_bias(st + IRS_VI_BETA_POW, beta1, beta2, step0, bc1, bc2);
    st[IRS_VI_STEP] = step0 + 1.0;
}

}  // namespace

extern "C" int irs_vi_step(const irs_sgld_config* cfg, const irs_sgld_buffers* b, const irs_vi_buffers* vi, void* stream) {
    if (!cfg || !b || !vi) return IRS_ERR_BAD_ARG;
    if (cfg->C != 2 || cfg->tau != 0.0 || cfg->hyper_mode != IRS_HYPER_REFERENCE || b->sigma != nullptr) return IRS_ERR_BAD_ARG;
    if (!vi->mu || !vi->log_var || !vi->u || !vi->eps_store || !vi->vi_state || !vi->partials || !vi->counter || !b->v || !b->grad_v)
        return IRS_ERR_BAD_ARG;
    for (int k = 0; k < 3; ++k)
        if (!vi->adam_m[k] || !vi->adam_v[k]) return IRS_ERR_BAD_ARG;
    if (!(vi->lr_mu >= 0.0) || !(vi->lr_log_var >= 0.0) || !(vi->lr_u >= 0.0) || !(vi->lr_decay >= 0.0)) return IRS_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const bool ffd = cfg->ffd_cps[0] > 0;
    const long long Vs = ffd ? (long long)cfg->ffd_grid[0] * cfg->ffd_grid[1] * cfg->ffd_grid[2]
                             : (long long)cfg->D * cfg->H * cfg->W;
    const bool vec = (Vs % 4 == 0);
    long long nb = (Vs / (vec ? 4 : 1) + 255) / 256;
    if (nb > 592) nb = 592;
    if (nb < 1) nb = 1;
    vi_sample_kernel<<<(unsigned)nb, 256, 0, st>>>(*vi, cfg->seed, b->v, Vs);
    IRS_LAUNCH_CHECK();
    IRS_TRY(irs_sgld_step_scaled(cfg, b, stream, 0.5));
    long long nu = (3 * Vs + 255) / 256;
    if (nu > 148 * 16) nu = 148 * 16;
    vi_update_kernel<<<(unsigned)nu, 256, 0, st>>>(*vi, b->grad_v, Vs);
    IRS_LAUNCH_CHECK();
    vi_advance_kernel<<<1, 32, 0, st>>>(vi->vi_state, vi->beta1, vi->beta2);
    return (int)cudaGetLastError();
}
