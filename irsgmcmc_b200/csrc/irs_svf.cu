// irs_svf.cu -- scaling and squaring of a stationary velocity field and its adjoint
// (reference utils/transformation.py:63-76 = 12 x F.grid_sample; backward = 12 x grid_sampler_3d_backward)
//
// Voxel-unit recurrence:
//   u_0 = v / 2^n ;  u_{k+1}(i) = u_k(i) + trilinear[u_k]( clamp(i + u_k(i)) )
// Adjoint of one step, g' = dL/du_{k+1}:
//   g(t) = g'(t)                                              direct
//        + inside(t) * sum_c g'_c(t) * grad trilinear[u_c](p(t))   position term (gather)
//        + sum_s  w(p(s) - t) * g'(s)                         interpolation transpose
// The transpose is evaluated as a GATHER over the sources s in a window of radius R = floor(max|u_k|) + 1 around t
// (every s with a non-zero hat weight lies inside it), so no atomics are issued -- ATen scatters 24 atomicAdds per
// voxel here.  Candidates are pruned axis by axis (a zero x-weight skips the y/z loads).  For R above
// gather_radius_max the exact scatter kernel below takes over (large deformations; still CUDA, no CPU path).
#include "irs_kernels.cuh"
#include "irs_bodies.cuh"

namespace {

__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
    atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void __launch_bounds__(256)
svf_step_fwd_kernel(const float* __restrict__ in, float in_scale, float* __restrict__ out,
                    float* __restrict__ maxabs, IrsDims d) {
    const long long V = d.V();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    float m = 0.f;
    if (i < V) m = irs_body_svf_fwd(in + (size_t)c * 3 * V, in_scale, out + (size_t)c * 3 * V, V, i, d);
    // block max -> one atomic per block (max is order independent: deterministic)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mm = sm[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mm = fmaxf(mm, sm[w]);
        if (mm > 0.f) atomic_max_nonneg(maxabs, mm);
    }
}

// one adjoint step; see the header comment.  `in` = raw input of the forward step (scaled by in_scale on the fly)
__global__ void __launch_bounds__(256)
svf_step_bwd_kernel(const float* __restrict__ in, float in_scale, const float* __restrict__ gp_all,
                    float* __restrict__ g_all, const float* __restrict__ maxabs, int radius_max, float out_scale,
                    IrsDims d) {
    const long long V = d.V();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const size_t off = (size_t)blockIdx.y * 3 * V;
    const int R = (int)floorf(__ldg(maxabs)) + 1;
    irs_body_svf_bwd(in + off, in_scale, gp_all + off, g_all + off, R <= radius_max ? R : -1, out_scale, V, i, d);
}

// exact scatter form of the interpolation transpose for steps whose displacement exceeds the gather window
__global__ void __launch_bounds__(256)
svf_step_bwd_scatter_kernel(const float* __restrict__ in, float in_scale, const float* __restrict__ gp_all,
                            float* __restrict__ g_all, const float* __restrict__ maxabs, int radius_max,
                            float out_scale, IrsDims d) {
    const int R = (int)floorf(__ldg(maxabs)) + 1;
    if (R <= radius_max) return;
    const long long V = d.V();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const size_t off = (size_t)blockIdx.y * 3 * V;
    float* g = g_all + off;
    irs_body_svf_bwd_scatter(in + off, in_scale, gp_all + off, out_scale, V, i, d,
                             [&](long long t, int ch, float val) { atomicAdd(g + (size_t)ch * V + t, val); });
}

__global__ void __launch_bounds__(256)
svf_outputs_kernel(const float* __restrict__ u_all, const float* __restrict__ lin_x, const float* __restrict__ lin_y,
                   const float* __restrict__ lin_z, float* __restrict__ T, float* __restrict__ disp, IrsDims d) {
    const long long V = d.V();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const int c = blockIdx.y;
    const float* u = u_all + (size_t)c * 3 * V;
    const int x = (int)(i % d.W), y = (int)((i / d.W) % d.H), z = (int)(i / ((long long)d.W * d.H));
    const float ux = u[i], uy = u[V + i], uz = u[2 * V + i];
    if (disp != nullptr) {
        float* o = disp + (size_t)c * 3 * V;
        o[i] = ux; o[V + i] = uy; o[2 * V + i] = uz;
    }
    if (T != nullptr) {
        // reference utils/util.py:418-429: channel idx scaled by 2/(shape[2+idx]-1) (sic), identity = fp32 linspace
        float* o = T + (size_t)c * 3 * V;
        o[i] = lin_x[x] + ux * (2.0f / (float)(d.D - 1));
        o[V + i] = lin_y[y] + uy * (2.0f / (float)(d.H - 1));
        o[2 * V + i] = lin_z[z] + uz * (2.0f / (float)(d.W - 1));
    }
}

}  // namespace

int irs_launch_svf_fwd(const float* v, float* hist, float* maxabs, int n_steps, int C, IrsDims d, cudaStream_t st) {
    const size_t F = (size_t)C * 3 * d.V();
    cudaError_t e = cudaMemsetAsync(maxabs, 0, sizeof(float) * n_steps, st);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)((d.V() + 255) / 256), C);
    const float scale0 = 1.0f / (float)(1 << n_steps);
    for (int k = 0; k < n_steps; ++k) {
        const float* in = (k == 0) ? v : hist + (size_t)(k - 1) * F;
        svf_step_fwd_kernel<<<grid, 256, 0, st>>>(in, k == 0 ? scale0 : 1.0f, hist + (size_t)k * F, maxabs + k, d);
    }
    return (int)cudaGetLastError();
}

int irs_launch_svf_bwd(const float* v, const float* hist, const float* maxabs, float* g_u, float* g_work, float* g_v,
                       int n_steps, int gather_radius_max, int C, IrsDims d, cudaStream_t st) {
    const size_t F = (size_t)C * 3 * d.V();
    dim3 grid((unsigned)((d.V() + 255) / 256), C);
    const float scale0 = 1.0f / (float)(1 << n_steps);
    // ping-pong between g_work and the caller's g_u buffer (g_u is only read by the first adjoint step)
    const float* gp = g_u;
    for (int k = n_steps - 1; k >= 0; --k) {
        const float* in = (k == 0) ? v : hist + (size_t)(k - 1) * F;
        float* out = (k == 0) ? g_v : (((n_steps - 1 - k) & 1) ? g_u : g_work);
        const float in_scale = (k == 0) ? scale0 : 1.0f;
        svf_step_bwd_kernel<<<grid, 256, 0, st>>>(in, in_scale, gp, out, maxabs + k, gather_radius_max, in_scale, d);
        svf_step_bwd_scatter_kernel<<<grid, 256, 0, st>>>(in, in_scale, gp, out, maxabs + k, gather_radius_max,
                                                          in_scale, d);
        gp = out;
    }
    return (int)cudaGetLastError();
}

extern "C" size_t irs_svf_hist_floats(int C, int D, int H, int W, int n_steps) {
    return (size_t)n_steps * C * 3 * D * H * W;
}

extern "C" int irs_svf_exp_fwd(const float* v, float* hist, float* maxabs, int n_steps, int C, int D, int H, int W,
                               void* stream) {
    IRS_CHECK_DIMS(C, D, H, W);
    if (!v || !hist || !maxabs || n_steps < 1 || n_steps > IRS_MAX_SVF_STEPS) return IRS_ERR_BAD_ARG;
    return irs_launch_svf_fwd(v, hist, maxabs, n_steps, C, IrsDims{D, H, W}, (cudaStream_t)stream);
}

extern "C" int irs_svf_outputs(const float* u, const float* lin_x, const float* lin_y, const float* lin_z, float* T,
                               float* disp, int C, int D, int H, int W, void* stream) {
    IRS_CHECK_DIMS(C, D, H, W);
    if (!u || (T && (!lin_x || !lin_y || !lin_z))) return IRS_ERR_BAD_ARG;
    IrsDims d{D, H, W};
    dim3 grid((unsigned)((d.V() + 255) / 256), C);
    svf_outputs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(u, lin_x, lin_y, lin_z, T, disp, d);
    return (int)cudaGetLastError();
}

extern "C" int irs_svf_exp_bwd(const float* v, const float* hist, const float* maxabs, float* g_u, float* g_work,
                               float* g_v, int n_steps, int gather_radius_max, int C, int D, int H, int W,
                               void* stream) {
    IRS_CHECK_DIMS(C, D, H, W);
    if (!v || !hist || !maxabs || !g_u || !g_work || !g_v || n_steps < 1 || n_steps > IRS_MAX_SVF_STEPS)
        return IRS_ERR_BAD_ARG;
    if (gather_radius_max < 0) return IRS_ERR_BAD_ARG;
    return irs_launch_svf_bwd(v, hist, maxabs, g_u, g_work, g_v, n_steps, gather_radius_max, C, IrsDims{D, H, W},
                              (cudaStream_t)stream);
}
