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

// large-displacement regime (R > radius_max): direct + position terms; the scatter kernel then adds the transpose
__global__ void __launch_bounds__(256)
svf_step_bwd_scatter_pre_kernel(const float* __restrict__ in, float in_scale, const float* __restrict__ gp_all,
                                float* __restrict__ g_all, const float* __restrict__ maxabs, int radius_max,
                                float out_scale, IrsDims d) {
    const int R = (int)floorf(__ldg(maxabs)) + 1;
    if (R <= radius_max) return;
    const long long V = d.V();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const size_t off = (size_t)blockIdx.y * 3 * V;
    irs_body_svf_bwd(in + off, in_scale, gp_all + off, g_all + off, -1, out_scale, V, i, d);
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

// ---------------------------------------------------------------------------------------------------------------------
// Tiled adjoint step.  A CTA owns a TX x TY column of targets and marches over source planes; the velocity planes
// s-R..s+R live in a shared-memory ring (halo R in x and y), the incoming gradient plane s in a second buffer.
// Every source voxel is visited once per (ox, oy) offset and deposits into the 2R+1 register accumulators of the
// target planes s-R..s+R, so a target voxel costs (2R+1)^2 pruned candidates instead of (2R+1)^3, all from shared
// memory; the position term gathers the same ring.  Radii 1 and 2 are compiled; larger ones fall back to the global
// gather below (or to the scatter kernel above gather_radius_max).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int BT_X = 32, BT_Y = 8;

template <int R>
__device__ __forceinline__ void svf_bwd_tile_body(const float* __restrict__ u, float in_scale,
                                                  const float* __restrict__ gp, float* __restrict__ g, float out_scale,
                                                  IrsDims d, int x0t, int y0t, int zs, int ze, float* smem) {
    constexpr int EX = BT_X + 2 * R, EY = BT_Y + 2 * R, PS = EX * EY, NP = 2 * R + 1;
    float* U = smem;               // [3][NP][EY][EX]  ring of velocity planes, slot = plane mod NP
    float* G = smem + 3 * NP * PS; // [3][EY][EX]      incoming gradient of the current source plane
    const long long V = d.V();
    const int tid = threadIdx.x, lx = tid % BT_X, ly = tid / BT_X;
    const int x = x0t + lx, y = y0t + ly;
    const bool active = x < d.W && y < d.H;
    const float xmax = (float)(d.W - 1), ymax = (float)(d.H - 1), zmax = (float)(d.D - 1);

    auto load_u_plane = [&](int pz) {
        if (pz < 0 || pz >= d.D) return;
        const int slot = pz % NP;
        for (int idx = tid; idx < PS; idx += BT_X * BT_Y) {
            const int ey = idx / EX, ex = idx - ey * EX;
            const int gx = x0t - R + ex, gy = y0t - R + ey;
            const bool ok = gx >= 0 && gx < d.W && gy >= 0 && gy < d.H;
            const long long gi = ((long long)pz * d.H + gy) * d.W + gx;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) U[(ch * NP + slot) * PS + idx] = ok ? __ldg(u + (size_t)ch * V + gi) * in_scale : 0.f;
        }
    };
    auto load_g_plane = [&](int pz) {
        if (pz < 0 || pz >= d.D) return;
        for (int idx = tid; idx < PS; idx += BT_X * BT_Y) {
            const int ey = idx / EX, ex = idx - ey * EX;
            const int gx = x0t - R + ex, gy = y0t - R + ey;
            const bool ok = gx >= 0 && gx < d.W && gy >= 0 && gy < d.H;
            const long long gi = ((long long)pz * d.H + gy) * d.W + gx;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) G[ch * PS + idx] = ok ? __ldg(gp + (size_t)ch * V + gi) : 0.f;
        }
    };

    float acc[NP][3];
#pragma unroll
    for (int i = 0; i < NP; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0.f;

    const int s_first = zs - R, s_last = ze - 1 + R;
    for (int pz = s_first - R; pz <= s_first + R; ++pz) load_u_plane(pz);
    load_g_plane(s_first);
    __syncthreads();

    for (int s = s_first; s <= s_last; ++s) {
        if (active && s >= 0 && s < d.D) {
            const int slot = s % NP;
            // ---- interpolation transpose: sources (x+ox, y+oy, s) deposit into targets (x, y, s-R..s+R) ----
#pragma unroll
            for (int oy = -R; oy <= R; ++oy) {
                const int sy = y + oy;
                if (sy < 0 || sy >= d.H) continue;
#pragma unroll
                for (int ox = -R; ox <= R; ++ox) {
                    const int sx = x + ox;
                    if (sx < 0 || sx >= d.W) continue;
                    const int li = (ly + R + oy) * EX + lx + R + ox;
                    const float wx = irs_hat(irs_clampf((float)sx + U[(0 * NP + slot) * PS + li], 0.f, xmax), x);
                    if (wx == 0.f) continue;
                    const float wy = irs_hat(irs_clampf((float)sy + U[(1 * NP + slot) * PS + li], 0.f, ymax), y);
                    if (wy == 0.f) continue;
                    const float pz = irs_clampf((float)s + U[(2 * NP + slot) * PS + li], 0.f, zmax);
                    const float wxy = wx * wy;
                    const float g0 = G[li], g1 = G[PS + li], g2 = G[2 * PS + li];
#pragma unroll
                    for (int dz = -R; dz <= R; ++dz) {
                        const float w = wxy * irs_hat(pz, s + dz);
                        acc[dz + R][0] += w * g0; acc[dz + R][1] += w * g1; acc[dz + R][2] += w * g2;
                    }
                }
            }
            // ---- direct + position term of target (x, y, s) ----
            if (s >= zs && s < ze) {
                const int lc = (ly + R) * EX + lx + R;
                const float g0 = G[lc], g1 = G[PS + lc], g2 = G[2 * PS + lc];
                float px = (float)x + U[(0 * NP + slot) * PS + lc];
                float py = (float)y + U[(1 * NP + slot) * PS + lc];
                float pz = (float)s + U[(2 * NP + slot) * PS + lc];
                const float mx = irs_inside(px, d.W), my = irs_inside(py, d.H), mz = irs_inside(pz, d.D);
                px = irs_clampf(px, 0.f, xmax); py = irs_clampf(py, 0.f, ymax); pz = irs_clampf(pz, 0.f, zmax);
                const float fx0 = floorf(px), fy0 = floorf(py), fz0 = floorf(pz);
                const int ix = (int)fx0, iy = (int)fy0, iz = (int)fz0;
                IrsCell cell;
                cell.fx = px - fx0; cell.fy = py - fy0; cell.fz = pz - fz0;
                cell.sx = (ix + 1 < d.W) ? 1 : 0;
                cell.sy = (iy + 1 < d.H) ? EX : 0;
                cell.sz = (iz + 1 < d.D) ? (((iz + 1) % NP) - (iz % NP)) * PS : 0;
                cell.i000 = (iz % NP) * PS + (iy - (y0t - R)) * EX + (ix - (x0t - R));
                float jx = 0.f, jy = 0.f, jz = 0.f, dx, dy, dz;
                irs_interp_grad(cell, [&](int k) { return U[k]; }, dx, dy, dz);
                jx += g0 * dx; jy += g0 * dy; jz += g0 * dz;
                irs_interp_grad(cell, [&](int k) { return U[NP * PS + k]; }, dx, dy, dz);
                jx += g1 * dx; jy += g1 * dy; jz += g1 * dz;
                irs_interp_grad(cell, [&](int k) { return U[2 * NP * PS + k]; }, dx, dy, dz);
                jx += g2 * dx; jy += g2 * dy; jz += g2 * dz;
                acc[R][0] += g0 + mx * jx; acc[R][1] += g1 + my * jy; acc[R][2] += g2 + mz * jz;
            }
        }
        // ---- target plane s-R is complete ----
        const int t = s - R;
        if (active && t >= zs && t < ze) {
            const long long gi = ((long long)t * d.H + y) * d.W + x;
            g[gi] = acc[0][0] * out_scale; g[V + gi] = acc[0][1] * out_scale; g[2 * V + gi] = acc[0][2] * out_scale;
        }
#pragma unroll
        for (int i = 0; i < NP - 1; ++i) { acc[i][0] = acc[i + 1][0]; acc[i][1] = acc[i + 1][1]; acc[i][2] = acc[i + 1][2]; }
        acc[NP - 1][0] = acc[NP - 1][1] = acc[NP - 1][2] = 0.f;
        __syncthreads();                 // everyone is done with plane s-R of the ring and with G
        load_u_plane(s + 1 + R);         // overwrites the slot of plane s-R
        load_g_plane(s + 1);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(BT_X * BT_Y)
svf_step_bwd_tile_kernel(const float* __restrict__ in, float in_scale, const float* __restrict__ gp_all,
                         float* __restrict__ g_all, const float* __restrict__ maxabs, int radius_max, float out_scale,
                         int seg_len, IrsDims d) {
    extern __shared__ float smem[];
    const int R = (int)floorf(__ldg(maxabs)) + 1;
    if (R > radius_max) return;  // the scatter kernel owns this step (after svf_step_bwd_kernel wrote the other terms)
    const long long V = d.V();
    const size_t off = (size_t)blockIdx.y * 3 * V;
    const int tiles_x = (d.W + BT_X - 1) / BT_X, tiles_y = (d.H + BT_Y - 1) / BT_Y;
    const int bx = blockIdx.x % tiles_x, by = (blockIdx.x / tiles_x) % tiles_y, bz = blockIdx.x / (tiles_x * tiles_y);
    const int x0t = bx * BT_X, y0t = by * BT_Y, zs = bz * seg_len, ze = min(zs + seg_len, d.D);
    if (R == 1) {
        svf_bwd_tile_body<1>(in + off, in_scale, gp_all + off, g_all + off, out_scale, d, x0t, y0t, zs, ze, smem);
    } else if (R == 2) {
        svf_bwd_tile_body<2>(in + off, in_scale, gp_all + off, g_all + off, out_scale, d, x0t, y0t, zs, ze, smem);
    } else {  // rare: wide gather straight from global memory
        const int x = x0t + (threadIdx.x % BT_X), y = y0t + (threadIdx.x / BT_X);
        if (x >= d.W || y >= d.H) return;
        for (int z = zs; z < ze; ++z)
            irs_body_svf_bwd(in + off, in_scale, gp_all + off, g_all + off, R, out_scale, V,
                             ((long long)z * d.H + y) * d.W + x, d);
    }
}

constexpr size_t svf_bwd_tile_smem(int R) {
    return sizeof(float) * (size_t)(3 * (2 * R + 1) + 3) * (BT_X + 2 * R) * (BT_Y + 2 * R);
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
    const float scale0 = 1.0f / (float)(1 << n_steps);
    static bool configured = false;
    const size_t smem = svf_bwd_tile_smem(2);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(svf_step_bwd_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const int tiles = ((d.W + BT_X - 1) / BT_X) * ((d.H + BT_Y - 1) / BT_Y);
    // z segments: long enough to amortise the 2R extra source planes, short enough to fill 148 SMs
    int seg_len = 32;
    while (seg_len > 8 && (long long)tiles * ((d.D + seg_len - 1) / seg_len) * C < 4 * 148) seg_len /= 2;
    const int nseg = (d.D + seg_len - 1) / seg_len;
    dim3 tgrid(tiles * nseg, C);
    dim3 vgrid((unsigned)((d.V() + 255) / 256), C);
    // ping-pong between g_work and the caller's g_u buffer (g_u is only read by the first adjoint step)
    const float* gp = g_u;
    for (int k = n_steps - 1; k >= 0; --k) {
        const float* in = (k == 0) ? v : hist + (size_t)(k - 1) * F;
        float* out = (k == 0) ? g_v : (((n_steps - 1 - k) & 1) ? g_u : g_work);
        const float in_scale = (k == 0) ? scale0 : 1.0f;
        svf_step_bwd_tile_kernel<<<tgrid, BT_X * BT_Y, smem, st>>>(in, in_scale, gp, out, maxabs + k, gather_radius_max,
                                                                  in_scale, seg_len, d);
        svf_step_bwd_scatter_pre_kernel<<<vgrid, 256, 0, st>>>(in, in_scale, gp, out, maxabs + k, gather_radius_max,
                                                               in_scale, d);
        svf_step_bwd_scatter_kernel<<<vgrid, 256, 0, st>>>(in, in_scale, gp, out, maxabs + k, gather_radius_max,
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
