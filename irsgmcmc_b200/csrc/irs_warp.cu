// irs_warp.cu -- trilinear / nearest-neighbour warps (reference utils/registration.py:17-32; jitter utils/util.py:44-53)
//
// HBM-bound gathers: one thread per output voxel, coalesced streaming reads of the grid, the gathered image goes
// through the read-only path (it is shared by all chains and stays L2-resident).
#include "irs_kernels.cuh"
#include "irs_bodies.cuh"

namespace {

// sampling position of voxel (x,y,z) from a normalised grid T (+ optional jitter), in ATen's operation order
__device__ __forceinline__ void position_from_T(const float* __restrict__ T, const float* __restrict__ jit, float alpha,
                                                long long V, long long i, IrsDims d, float& px, float& py, float& pz) {
    float gx = T[i], gy = T[V + i], gz = T[2 * V + i];
    if (jit != nullptr) {
        // reference utils/util.py:44-45,52-53,418-429: T + ((-2 alpha U + alpha) * 2 / (dims[idx] - 1)); channel idx uses
        // tensor dim 2+idx (sic)
        gx += irs_jitter_normalised(jit[i], alpha, d.D);
        gy += irs_jitter_normalised(jit[V + i], alpha, d.H);
        gz += irs_jitter_normalised(jit[2 * V + i], alpha, d.W);
    }
    px = irs_unnormalise(gx, d.W);
    py = irs_unnormalise(gy, d.H);
    pz = irs_unnormalise(gz, d.D);
}

__global__ void __launch_bounds__(256)
warp_fwd_kernel(const float* __restrict__ img, long long img_cs, const float* __restrict__ T,
                const float* __restrict__ jit, float alpha, float* __restrict__ out, IrsDims d) {
    const long long V = d.V();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const int c = blockIdx.y;
    const float* Tc = T + (size_t)c * 3 * V;
    const float* jc = jit ? jit + (size_t)c * 3 * V : nullptr;
    const float* im = img + (size_t)c * img_cs;
    float px, py, pz;
    position_from_T(Tc, jc, alpha, V, i, d, px, py, pz);
    out[(size_t)c * V + i] = irs_body_warp_fwd(im, px, py, pz, d);
}

__global__ void __launch_bounds__(256)
warp_bwd_grid_kernel(const float* __restrict__ img, long long img_cs, const float* __restrict__ T,
                     const float* __restrict__ jit, float alpha, const float* __restrict__ g_out,
                     float* __restrict__ g_T, IrsDims d) {
    const long long V = d.V();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const int c = blockIdx.y;
    const float* Tc = T + (size_t)c * 3 * V;
    const float* jc = jit ? jit + (size_t)c * 3 * V : nullptr;
    const float* im = img + (size_t)c * img_cs;
    float px, py, pz;
    position_from_T(Tc, jc, alpha, V, i, d, px, py, pz);
    float gx, gy, gz;
    irs_body_warp_grad(im, px, py, pz, d, g_out[(size_t)c * V + i], 0.5f * (float)(d.W - 1), 0.5f * (float)(d.H - 1),
                       0.5f * (float)(d.D - 1), gx, gy, gz);
    float* g = g_T + (size_t)c * 3 * V;
    g[i] = gx; g[V + i] = gy; g[2 * V + i] = gz;
}

template <typename T_>
__global__ void __launch_bounds__(256)
warp_nearest_kernel(const T_* __restrict__ seg, long long seg_cs, const float* __restrict__ T, T_* __restrict__ out,
                    IrsDims d) {
    const long long V = d.V();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const int c = blockIdx.y;
    const float* Tc = T + (size_t)c * 3 * V;
    out[(size_t)c * V + i] = seg[(size_t)c * seg_cs + irs_body_nearest_index(Tc, V, i, d)];
}


// ---- voxel-unit variants used inside the fused step: position = index + u (+ jitter in voxels) ----------------------
// 32-bit index arithmetic throughout (3 V < 2^31 is checked at the ABI): 64-bit divisions cost more than the gather.
__device__ __forceinline__ void position_from_u(const float* __restrict__ u, const IrsRng& jit, float alpha,
                                                int use_jitter, int V, int i, int c, IrsDims d, float& px, float& py,
                                                float& pz) {
    const unsigned q = (unsigned)i / (unsigned)d.W, x = (unsigned)i - q * (unsigned)d.W;
    const unsigned z = q / (unsigned)d.H, y = q - z * (unsigned)d.H;
    px = (float)x + u[i];
    py = (float)y + u[V + i];
    pz = (float)z + u[2 * V + i];
    if (use_jitter) {
        float r[3];
        if (jit.explicit_values != nullptr) {
            const float* j = jit.explicit_values + (size_t)c * 3 * V;
            r[0] = j[i]; r[1] = j[V + i]; r[2] = j[2 * V + i];
        } else {
            const unsigned long long it = jit.iter_ptr ? (unsigned long long)(*jit.iter_ptr) : jit.iter;
            irs_uniform3(jit.seed, (uint32_t)i, (uint32_t)(jit.chain0 + c), it, r);
        }
        px += alpha - 2.0f * alpha * r[0];
        py += alpha - 2.0f * alpha * r[1];
        pz += alpha - 2.0f * alpha * r[2];
    }
}

// GRAD: also leaves d out / d position (zero where the position sits on / outside the border) in grad (C,3,V): the adjoint of
// the warp is then a multiplication by the incoming gradient -- an epilogue of the residual map's adjoint -- instead of a second
// gather kernel that re-derives the position, re-generates the jitter and re-gathers the eight corners.
template <bool GRAD>
__global__ void __launch_bounds__(256)
warp_vox_fwd_kernel(const float* __restrict__ img, const float* __restrict__ u, IrsRng jit, float alpha, int use_jitter,
                    float* __restrict__ out, float* __restrict__ grad, IrsDims d) {
    const int V = (int)d.V();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const int c = blockIdx.y;
    float px, py, pz;
    position_from_u(u + (size_t)c * 3 * V, jit, alpha, use_jitter, V, i, c, d, px, py, pz);
    if (!GRAD) {
        out[(size_t)c * V + i] = irs_body_warp_fwd(img, px, py, pz, d);
    } else {
        const float mx = irs_inside(px, d.W), my = irs_inside(py, d.H), mz = irs_inside(pz, d.D);
        px = irs_clampf(px, 0.f, (float)(d.W - 1));
        py = irs_clampf(py, 0.f, (float)(d.H - 1));
        pz = irs_clampf(pz, 0.f, (float)(d.D - 1));
        const IrsCell cell = irs_cell(px, py, pz, d);
        float dx, dy, dz;
        out[(size_t)c * V + i] = irs_interp_grad(cell, [&](int k) { return __ldg(img + k); }, dx, dy, dz);
        float* g = grad + (size_t)c * 3 * V;
        g[i] = dx * mx; g[V + i] = dy * my; g[2 * V + i] = dz * mz;
    }
}

// g_u = (sign g_out) * grad, in place over grad (SSD: the residual map's adjoint is a sign)
__global__ void __launch_bounds__(256)
warp_apply_grad_kernel(const float* __restrict__ g_out, float sign, float* __restrict__ grad, IrsDims d) {
    const int V = (int)d.V();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const int c = blockIdx.y;
    const float go = sign * g_out[(size_t)c * V + i];
    float* g = grad + (size_t)c * 3 * V;
    g[i] = go * g[i]; g[V + i] = go * g[V + i]; g[2 * V + i] = go * g[2 * V + i];
}

__global__ void __launch_bounds__(256)
warp_vox_bwd_kernel(const float* __restrict__ img, const float* __restrict__ u, IrsRng jit, float alpha, int use_jitter,
                    const float* __restrict__ g_out, float g_sign, float* __restrict__ g_u, IrsDims d) {
    const int V = (int)d.V();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const int c = blockIdx.y;
    float px, py, pz;
    position_from_u(u + (size_t)c * 3 * V, jit, alpha, use_jitter, V, i, c, d, px, py, pz);
    float gx, gy, gz;
    irs_body_warp_grad(img, px, py, pz, d, g_sign * g_out[(size_t)c * V + i], 1.f, 1.f, 1.f, gx, gy, gz);
    float* g = g_u + (size_t)c * 3 * V;
    g[i] = gx; g[V + i] = gy; g[2 * V + i] = gz;
}

}  // namespace

extern "C" int irs_warp3d_fwd(const float* img, long long img_cs, const float* T, const float* jit, float alpha,
                              float* out, int C, int D, int H, int W, void* stream) {
    IRS_CHECK_DIMS(C, D, H, W);
    if (!img || !T || !out) return IRS_ERR_BAD_ARG;
    IrsDims d{D, H, W};
    dim3 grid((unsigned)((d.V() + 255) / 256), C);
    warp_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, img_cs, T, jit, alpha, out, d);
    return (int)cudaGetLastError();
}

extern "C" int irs_warp3d_bwd_grid(const float* img, long long img_cs, const float* T, const float* jit, float alpha,
                                   const float* g_out, float* g_T, int C, int D, int H, int W, void* stream) {
    IRS_CHECK_DIMS(C, D, H, W);
    if (!img || !T || !g_out || !g_T) return IRS_ERR_BAD_ARG;
    IrsDims d{D, H, W};
    dim3 grid((unsigned)((d.V() + 255) / 256), C);
    warp_bwd_grid_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, img_cs, T, jit, alpha, g_out, g_T, d);
    return (int)cudaGetLastError();
}

extern "C" int irs_warp3d_nearest_i16(const short* seg, long long seg_cs, const float* T, short* out, int C, int D,
                                      int H, int W, void* stream) {
    IRS_CHECK_DIMS(C, D, H, W);
    if (!seg || !T || !out) return IRS_ERR_BAD_ARG;
    IrsDims d{D, H, W};
    dim3 grid((unsigned)((d.V() + 255) / 256), C);
    warp_nearest_kernel<short><<<grid, 256, 0, (cudaStream_t)stream>>>(seg, seg_cs, T, out, d);
    return (int)cudaGetLastError();
}

extern "C" int irs_warp3d_nearest_u8(const unsigned char* seg, long long seg_cs, const float* T, unsigned char* out,
                                     int C, int D, int H, int W, void* stream) {
    IRS_CHECK_DIMS(C, D, H, W);
    if (!seg || !T || !out) return IRS_ERR_BAD_ARG;
    IrsDims d{D, H, W};
    dim3 grid((unsigned)((d.V() + 255) / 256), C);
    warp_nearest_kernel<unsigned char><<<grid, 256, 0, (cudaStream_t)stream>>>(seg, seg_cs, T, out, d);
    return (int)cudaGetLastError();
}

int irs_launch_warp_vox_fwd(const float* img, const float* u, IrsRng jit, float alpha, int use_jitter, float* out, int C,
                            IrsDims d, cudaStream_t st, float* grad) {
    dim3 grid((unsigned)((d.V() + 255) / 256), C);
    if (grad != nullptr) warp_vox_fwd_kernel<true><<<grid, 256, 0, st>>>(img, u, jit, alpha, use_jitter, out, grad, d);
    else warp_vox_fwd_kernel<false><<<grid, 256, 0, st>>>(img, u, jit, alpha, use_jitter, out, nullptr, d);
    return (int)cudaGetLastError();
}

int irs_launch_warp_apply_grad(const float* g_out, float sign, float* grad, int C, IrsDims d, cudaStream_t st) {
    dim3 grid((unsigned)((d.V() + 255) / 256), C);
    warp_apply_grad_kernel<<<grid, 256, 0, st>>>(g_out, sign, grad, d);
    return (int)cudaGetLastError();
}

int irs_launch_warp_vox_bwd(const float* img, const float* u, IrsRng jit, float alpha, int use_jitter,
                            const float* g_out, float g_sign, float* g_u, int C, IrsDims d, cudaStream_t st) {
    dim3 grid((unsigned)((d.V() + 255) / 256), C);
    warp_vox_bwd_kernel<<<grid, 256, 0, st>>>(img, u, jit, alpha, use_jitter, g_out, g_sign, g_u, d);
    return (int)cudaGetLastError();
}
