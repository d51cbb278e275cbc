// irs_common.cuh -- per-voxel arithmetic shared by every kernel of the SGLD registration step.
//
// Everything numerical lives here as __host__ __device__ inline functions so that tests/host_emul.cu can run the very
// same arithmetic on the CPU against the oracle (test infrastructure only -- the product has no CPU path).
//
// Conventions: volumes are contiguous (D,H,W), vector fields planar (3,D,H,W) with channel 0 = x
// (W axis), 1 = y (H axis), 2 = z (D axis).  Displacements are carried in VOXEL units inside the library; the
// reference's normalised [-1,1] units appear only at the op boundary (SURVEY Appendix A.6 shows the two are the same
// function).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define IRS_HD __host__ __device__ __forceinline__

#define IRS_OK 0
#define IRS_ERR_BAD_ARG (-1)
#define IRS_ERR_UNSUPPORTED (-2)
#define IRS_ERR_WORKSPACE (-3)

#define IRS_MAX_K 8        // GMM components
#define IRS_MAX_TAPS 15    // Sobolev kernel width (s <= 7)
#define IRS_MAX_SVF_STEPS 16

struct IrsDims {
    int D, H, W;
    IRS_HD long long V() const { return (long long)D * H * W; }
};

IRS_HD int irs_clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
IRS_HD float irs_clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(x, lo)); }

// ---------------------------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al. 2011).  Keyed by (seed), counted by (voxel, chain, iteration, stream),
// so a value can be regenerated anywhere (halos, backward pass) without state.
// ---------------------------------------------------------------------------------------------------------------------
struct IrsU4 { uint32_t x, y, z, w; };

IRS_HD uint32_t irs_mulhi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

IRS_HD IrsU4 irs_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = irs_mulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = irs_mulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    IrsU4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

#define IRS_STREAM_LANGEVIN 0x4c414e47u  // 'LANG'
#define IRS_STREAM_JITTER   0x4a495454u  // 'JITT'

// uniform in [0,1) with 24 random bits (what torch.rand produces for fp32)
IRS_HD float irs_u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// three independent N(0,1) for the three components of one voxel (Box-Muller on (0,1] x [0,1))
IRS_HD void irs_normal3(uint64_t seed, uint32_t voxel, uint32_t chain, uint64_t iter, float* e) {
    IrsU4 r = irs_philox(voxel, chain, (uint32_t)iter, (uint32_t)(iter >> 32) ^ IRS_STREAM_LANGEVIN,
                         (uint32_t)seed, (uint32_t)(seed >> 32));
    float u1 = ((float)(r.x >> 8) + 1.0f) * (1.0f / 16777216.0f);
    float u2 = ((float)(r.z >> 8) + 1.0f) * (1.0f / 16777216.0f);
    float s1, c1, s2, c2;
#ifdef __CUDA_ARCH__
    // Device: the hardware approximations (lg2 / sin / cos with 2^-21-level absolute error).  The generator is 233 instructions per
    // voxel with the accurate library calls -- as much as a squaring step's adjoint -- and 40 % fewer with these; a sampler's
    // N(0,1) needs the distribution, not the last bit (tests: moments, correlations, the posterior statistics against the
    // oracle sampler).  The angle pi (2 u - 1) in [-pi, pi) stays inside the range the fast sine is accurate in; shifting the
    // uniform angle by pi flips the sign of sine and cosine together, which leaves the distribution unchanged.
    float r1 = sqrtf(-2.0f * __logf(u1)), r2 = sqrtf(-2.0f * __logf(u2));
    __sincosf(3.14159265358979f * (2.0f * irs_u01(r.y) - 1.0f), &s1, &c1);
    c2 = __cosf(3.14159265358979f * (2.0f * irs_u01(r.w) - 1.0f));
    s2 = 0.f;
#else
    float r1 = sqrtf(-2.0f * logf(u1)), r2 = sqrtf(-2.0f * logf(u2));
    s1 = sinf(6.283185307179586f * irs_u01(r.y)); c1 = cosf(6.283185307179586f * irs_u01(r.y));
    s2 = sinf(6.283185307179586f * irs_u01(r.w)); c2 = cosf(6.283185307179586f * irs_u01(r.w));
#endif
    e[0] = r1 * c1; e[1] = r1 * s1; e[2] = r2 * c2;
    (void)s2;
}

// three independent U[0,1) for the jitter of one voxel
IRS_HD void irs_uniform3(uint64_t seed, uint32_t voxel, uint32_t chain, uint64_t iter, float* u) {
    IrsU4 r = irs_philox(voxel, chain, (uint32_t)iter, (uint32_t)(iter >> 32) ^ IRS_STREAM_JITTER,
                         (uint32_t)seed, (uint32_t)(seed >> 32));
    u[0] = irs_u01(r.x); u[1] = irs_u01(r.y); u[2] = irs_u01(r.z);
}

// ---------------------------------------------------------------------------------------------------------------------
// trilinear sampling with border clamp, align_corners=True  (reference utils/registration.py:29-30 -> ATen
// GridSampler.cuh:149-218 of torch 2.11; clip rules :53-81)
// ---------------------------------------------------------------------------------------------------------------------

// ATen grid_sampler_unnormalize, align_corners: ((g + 1) / 2) * (n - 1)  -- operation order kept, no contraction
IRS_HD float irs_unnormalise(float g, int n) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(n - 1));
#else
    volatile float a = g + 1.0f; volatile float b = a / 2; return b * (float)(n - 1);
#endif
}

struct IrsCell {
    int i000;            // linear index of the (x0,y0,z0) corner
    int sx, sy, sz;      // strides to the +1 corner along each axis (0 at the far border: that corner has weight 0)
    float fx, fy, fz;    // fractional position inside the cell
};

// px,py,pz must already be clamped to [0,n-1]
IRS_HD IrsCell irs_cell(float px, float py, float pz, IrsDims d) {
    IrsCell c;
    float x0 = floorf(px), y0 = floorf(py), z0 = floorf(pz);
    c.fx = px - x0; c.fy = py - y0; c.fz = pz - z0;
    int ix = (int)x0, iy = (int)y0, iz = (int)z0;
    c.sx = (ix + 1 < d.W) ? 1 : 0;
    c.sy = (iy + 1 < d.H) ? d.W : 0;
    c.sz = (iz + 1 < d.D) ? d.W * d.H : 0;
    c.i000 = (iz * d.H + iy) * d.W + ix;
    return c;
}

template <typename LD>
IRS_HD float irs_interp(const IrsCell& c, LD ld) {
    float v000 = ld(c.i000), v001 = ld(c.i000 + c.sx);
    float v010 = ld(c.i000 + c.sy), v011 = ld(c.i000 + c.sy + c.sx);
    float v100 = ld(c.i000 + c.sz), v101 = ld(c.i000 + c.sz + c.sx);
    float v110 = ld(c.i000 + c.sz + c.sy), v111 = ld(c.i000 + c.sz + c.sy + c.sx);
    float a00 = v000 + c.fx * (v001 - v000), a01 = v010 + c.fx * (v011 - v010);
    float a10 = v100 + c.fx * (v101 - v100), a11 = v110 + c.fx * (v111 - v110);
    float b0 = a00 + c.fy * (a01 - a00), b1 = a10 + c.fy * (a11 - a10);
    return b0 + c.fz * (b1 - b0);
}

// value and the derivative w.r.t. the (x,y,z) sampling position (the caller zeroes components whose coordinate sits
// on/outside the border: GridSampler.cuh:62-81)
template <typename LD>
IRS_HD float irs_interp_grad(const IrsCell& c, LD ld, float& gx, float& gy, float& gz) {
    float v000 = ld(c.i000), v001 = ld(c.i000 + c.sx);
    float v010 = ld(c.i000 + c.sy), v011 = ld(c.i000 + c.sy + c.sx);
    float v100 = ld(c.i000 + c.sz), v101 = ld(c.i000 + c.sz + c.sx);
    float v110 = ld(c.i000 + c.sz + c.sy), v111 = ld(c.i000 + c.sz + c.sy + c.sx);
    float d00 = v001 - v000, d01 = v011 - v010, d10 = v101 - v100, d11 = v111 - v110;
    float a00 = v000 + c.fx * d00, a01 = v010 + c.fx * d01, a10 = v100 + c.fx * d10, a11 = v110 + c.fx * d11;
    float e0 = a01 - a00, e1 = a11 - a10;
    float b0 = a00 + c.fy * e0, b1 = a10 + c.fy * e1;
    float dx0 = d00 + c.fy * (d01 - d00), dx1 = d10 + c.fy * (d11 - d10);
    gx = dx0 + c.fz * (dx1 - dx0);
    gy = e0 + c.fz * (e1 - e0);
    gz = b1 - b0;
    return b0 + c.fz * (b1 - b0);
}

// 1 where ATen propagates a gradient through the border clip: strictly inside (0, n-1)
IRS_HD float irs_inside(float p, int n) { return (p > 0.0f && p < (float)(n - 1)) ? 1.0f : 0.0f; }

// weight with which a sample at (clamped) position p deposits onto grid node t: the transpose of linear interpolation
IRS_HD float irs_hat(float p, int t) { return fmaxf(0.0f, 1.0f - fabsf(p - (float)t)); }

// nearest-neighbour source index, ATen nearest + border + align_corners: unnormalise, clip, nearbyint (half to even)
IRS_HD int irs_nearest_coord(float g, int n) {
    float p = irs_unnormalise(g, n);
    p = fminf((float)(n - 1), fmaxf(p, 0.0f));
    return (int)nearbyintf(p);
}

// ---------------------------------------------------------------------------------------------------------------------
// Gaussian mixture of K zero-mean components evaluated at one residual (reference model/loss.py:87-93)
// table per component: lw[k] = log pi_k - log sigma_k,  prec[k] = exp(-2 log sigma_k)
// ---------------------------------------------------------------------------------------------------------------------
struct IrsGmm {
    int K;
    float lw[IRS_MAX_K];
    float prec[IRS_MAX_K];
};

#define IRS_LOG_SQRT_2PI 0.9189385332046727f

// returns log pdf(z); rho[k] = responsibilities; wprec = sum_k rho_k prec_k  (so dNLL/dz = z*wprec, VD residual = z^2*wprec)
// KK = compile-time bound on the component loop (g.K <= KK): the kernels instantiate KK = 4 for the usual mixture so that
// no issue slots go to predicated-off components; same operations in the same order for every KK >= g.K.
template <int KK>
IRS_HD float irs_gmm_eval_t(const IrsGmm& g, float z, float* rho, float& wprec) {
    float e[KK];
    float hz2 = 0.5f * z * z, m = -INFINITY;
#pragma unroll
    for (int k = 0; k < KK; ++k) if (k < g.K) { e[k] = g.lw[k] - hz2 * g.prec[k]; m = fmaxf(m, e[k]); }
    float S = 0.0f;
#pragma unroll
    for (int k = 0; k < KK; ++k) if (k < g.K) { e[k] = expf(e[k] - m); S += e[k]; }
    float inv = 1.0f / S;
    wprec = 0.0f;
#pragma unroll
    for (int k = 0; k < KK; ++k) if (k < g.K) { rho[k] = e[k] * inv; wprec += rho[k] * g.prec[k]; }
    return m + logf(S) - IRS_LOG_SQRT_2PI;
}
IRS_HD float irs_gmm_eval(const IrsGmm& g, float z, float* rho, float& wprec) {
    return irs_gmm_eval_t<IRS_MAX_K>(g, z, rho, wprec);
}

// precision-weighted squared residual used by virtual decimation (reference utils/util.py:330-347, closed form)
IRS_HD float irs_gmm_vd_residual(const IrsGmm& g, float z) {
    float rho[IRS_MAX_K], wp;
    irs_gmm_eval(g, z, rho, wp);
    return z * z * wp;
}

// ---------------------------------------------------------------------------------------------------------------------
// regulariser: forward differences with the last difference duplicated (reference utils/diff_op.py:83-85)
//   energy along one axis  = sum_{i<=n-3} d_i^2 + 2 d_{n-2}^2 ;  d/dv_j = 2 (w_{j-1} d_{j-1} - w_j d_j), w_{n-2} = 2
// ---------------------------------------------------------------------------------------------------------------------
// contribution of axis position j (0..n-1) to the energy, given v[j] and v[j+1] (vp ignored when j == n-1)
IRS_HD float irs_diff_energy(float vj, float vp, int j, int n) {
    if (j >= n - 1) return 0.0f;
    float d = vp - vj;
    return (j == n - 2 ? 2.0f : 1.0f) * d * d;
}

// d energy / d v_j along one axis given the neighbours (vm = v[j-1], vp = v[j+1]; ignored when out of range)
IRS_HD float irs_diff_energy_grad(float vm, float vj, float vp, int j, int n) {
    float g = 0.0f;
    if (j >= 1) g += ((j - 1 == n - 2) ? 2.0f : 1.0f) * (vj - vm);
    if (j <= n - 2) g -= ((j == n - 2) ? 2.0f : 1.0f) * (vp - vj);
    return 2.0f * g;
}

// ---------------------------------------------------------------------------------------------------------------------
// device-side reductions
// ---------------------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float irs_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double irs_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum NV per-thread values over the block; valid in thread 0.  `sh` must hold NV * 32 doubles.
template <int NV>
__device__ __forceinline__ void irs_block_sum(const float* vals, double* out, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float s = irs_warp_sum(vals[i]);
        if (lane == 0) sh[i * 32 + warp] = (double)s;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = lane < nwarps ? sh[i * 32 + lane] : 0.0;
            s = irs_warp_sum(s);
            if (lane == 0) out[i] = s;
        }
    }
    __syncthreads();
}

// Deterministic grid reduction: every block stores its NV partial sums, the last block to arrive adds them up in a
// fixed order.  Returns true (in ALL threads of that last block) when `total` (shared memory, NV doubles) is valid.
// `partials` holds NV * gridDim.x doubles laid out [value][block]; `counter` is a zero-initialised uint that the last
// block resets.  Only the first `n_used` values are reduced.
template <int NV>
__device__ __forceinline__ bool irs_grid_sum(const double* block_vals, double* partials, unsigned int* counter,
                                             double* total, int n_used = NV) {
    __shared__ bool is_last;
    const unsigned int G = gridDim.x;
    if (threadIdx.x == 0) {
        for (int i = 0; i < n_used; ++i) partials[(size_t)i * G + blockIdx.x] = block_vals[i];
        __threadfence();
        unsigned int ticket = atomicAdd(counter, 1u);
        is_last = (ticket == G - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    // one warp per value, lanes stride over the blocks with four independent accumulators (fixed order -> deterministic)
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int i = threadIdx.x >> 5; i < n_used; i += nwarps) {
        const double* p = partials + (size_t)i * G;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        unsigned int b = lane;
        for (; b + 96 < G; b += 128) { s0 += p[b]; s1 += p[b + 32]; s2 += p[b + 64]; s3 += p[b + 96]; }
        for (; b < G; b += 32) s0 += p[b];
        double s = irs_warp_sum((s0 + s1) + (s2 + s3));
        if (lane == 0) total[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) *counter = 0u;
    return true;
}
#endif  // __CUDACC__
