// irs_ffd.cu -- cubic B-spline free-form deformation: control-point velocities -> dense velocity field and its adjoint
// (SURVEY.md section 8f, N3).  Replaces Cubic_B_spline_FFD_3D.forward / conv1D(transpose=True) and their autograd,
// reference utils/transformation.py:106-152.
//
// The tensor product is evaluated like the reference does, one axis at a time (z, y, x), so the two intermediates live on
// the coarse grid along the axes not yet expanded: at cps 4 and 128^3 they are 0.6 MB and 6.9 MB per chain next to the
// 25 MB result, i.e. the op moves about 18 B per voxel instead of the 12 B minimum, with 4 multiply-adds per output
// element and pass.  The adjoint runs the same three passes in reverse as gathers over each control point's support
// (4 s - 1 elements): deterministic, no atomics.  Arithmetic: csrc/irs_ffd_body.cuh.
#include <stdlib.h>

#include "irs_ffd_body.cuh"
#include "irs_kernels.cuh"

namespace {

template <int VEC, bool ADJOINT>
__global__ void __launch_bounds__(256)
ffd_axis_kernel(const float* __restrict__ in, float* __restrict__ out, unsigned groups, int g, int n, unsigned inner,
                const __grid_constant__ IrsFfdAxis ax) {
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < groups; i += stride) {
        float r[VEC];
        irs_body_ffd_axis_group<VEC, ADJOINT>(in, i, g, n, inner, ax, r);
        if constexpr (VEC == 4) reinterpret_cast<float4*>(out)[i] = make_float4(r[0], r[1], r[2], r[3]);
        else out[i] = r[0];
    }
}

// ---- the contiguous axis (inner == 1): the pass that touches the dense field ----------------------------------------------
// Forward: a thread owns ONE coordinate x for the whole launch (the grid is sized so that the thread count is a multiple
// of n) and marches over rows: its four control-point offsets and weights are computed once and stay in registers, an
// output is four loads, four FMAs and a coalesced store.
__global__ void __launch_bounds__(256)
ffd_fwd_rows_kernel(const float* __restrict__ cp, float* __restrict__ out, unsigned rows, int g, int n,
                    const __grid_constant__ IrsFfdAxis ax) {
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned row0 = gid / (unsigned)n, x = gid - row0 * (unsigned)n;
    const unsigned row_step = gridDim.x * blockDim.x / (unsigned)n;   // exact by construction of the grid
    const IrsFfdEntry e = irs_ffd_entry((int)x, g, ax);
    for (unsigned row = row0; row < rows; row += row_step)
        out[(size_t)row * n + x] = irs_body_ffd_axis_fwd_tab(cp + (size_t)row * g, 1u, e);
}

// Adjoint: a block stages `rb` dense rows (contiguous in memory, coalesced loads) in shared memory with the skew of
// irs_ffd_body.cuh, then every thread gathers control points from the staged rows without bank conflicts.
__global__ void __launch_bounds__(256)
ffd_bwd_rows_kernel(const float* __restrict__ gd, float* __restrict__ out, unsigned rows, int rb, int g, int n,
                    const __grid_constant__ IrsFfdAxis ax) {
    extern __shared__ float s_rows[];
    const IrsFfdSkew sk = irs_ffd_make_skew(ax.s);
    const int pitch = irs_ffd_row_pitch(n, sk);
    const unsigned step_x = blockDim.x % (unsigned)n, step_r = blockDim.x / (unsigned)n;
    for (unsigned r0 = blockIdx.x * (unsigned)rb; r0 < rows; r0 += gridDim.x * (unsigned)rb) {
        const unsigned here = rows - r0 < (unsigned)rb ? rows - r0 : (unsigned)rb;
        const float* src = gd + (size_t)r0 * n;
        unsigned r = threadIdx.x / (unsigned)n, x = threadIdx.x - r * (unsigned)n;
        for (unsigned e = threadIdx.x; e < here * (unsigned)n; e += blockDim.x) {
            s_rows[r * pitch + irs_ffd_skew((int)x, sk)] = __ldg(src + e);
            x += step_x;
            r += step_r;
            if (x >= (unsigned)n) { x -= (unsigned)n; ++r; }
        }
        __syncthreads();
        float* dst = out + (size_t)r0 * g;
        for (unsigned e = threadIdx.x; e < here * (unsigned)g; e += blockDim.x) {
            const unsigned rr = e / (unsigned)g, i = e - rr * (unsigned)g;
            dst[e] = irs_body_ffd_axis_bwd_row(s_rows + rr * pitch, sk, (int)i, n, ax);
        }
        __syncthreads();
    }
}

int make_axis(IrsFfdAxis& ax, const float* kernel_host, int s, int off) {
    if (!kernel_host || s < 1 || s > IRS_FFD_MAX_STRIDE || off < 0) return IRS_ERR_BAD_ARG;
    ax.s = s;
    ax.off = off;
    for (int j = 0; j <= IRS_FFD_MAX_KERNEL; ++j) ax.k[j] = j < 4 * s - 1 ? kernel_host[j] : 0.f;
    return IRS_OK;
}

// one axis pass; `adjoint` maps (outer, n, inner) -> (outer, g, inner), otherwise (outer, g, inner) -> (outer, n, inner)
int launch_axis(const float* in, float* out, bool adjoint, long long outer, int g, int n, long long inner,
                const IrsFfdAxis& ax, cudaStream_t st) {
    // every dense element must exist in the un-cropped result of the transposed convolution: (g - 1) s + 1 elements
    if (g < 1 || n < 1 || outer < 1 || inner < 1 || (long long)ax.off + n > (long long)(g - 1) * ax.s + 1)
        return IRS_ERR_BAD_ARG;
    // 32-bit flat indices inside a launch: larger arrays go in slices of whole `outer` rows (rows are independent)
    const long long row_in = (long long)(adjoint ? n : g) * inner, row_out = (long long)(adjoint ? g : n) * inner;
    if (row_in >= (1ll << 31) || row_out >= (1ll << 31)) return IRS_ERR_UNSUPPORTED;
    const long long rows_per_launch = (1ll << 31) / (row_in > row_out ? row_in : row_out);
    // development switch, read at every launch so that a test can compare the two paths: IRS_FFD_GENERIC=1 keeps the
    // generic kernel for every pass
    const char* gen_env = getenv("IRS_FFD_GENERIC");
    const bool generic_only = gen_env != nullptr && atoi(gen_env) == 1;
    for (long long o0 = 0; o0 < outer; o0 += rows_per_launch) {
        const long long rows = outer - o0 < rows_per_launch ? outer - o0 : rows_per_launch;
        const float* src = in + o0 * row_in;
        float* dst = out + o0 * row_out;
        const long long total = rows * row_out;
        if (inner == 1 && !generic_only && !adjoint && n <= 148 * 8 * 256) {
            // thread count = a multiple of n (so every thread keeps its x) close to 8 resident CTAs per SM
            long long per = n, blocks = 1;
            while ((per & 255) != 0) per <<= 1;            // lcm(n, 256) = n * 2^k
            const long long unit = per / 256;              // blocks per unit of lcm(n, 256) threads
            blocks = (total + per - 1) / per * unit;       // enough threads for one element each ...
            const long long cap = (148 * 8 / unit) * unit; // ... capped near 1184 blocks, a multiple of the unit
            if (cap >= unit && blocks > cap) blocks = cap;
            if (blocks >= unit && blocks <= 0x7fffffffll) {
                ffd_fwd_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, dst, (unsigned)rows, g, n, ax);
                IRS_LAUNCH_CHECK();
                continue;
            }
        }
        if (inner == 1 && !generic_only && adjoint) {
            const IrsFfdSkew sk = irs_ffd_make_skew(ax.s);
            const long long pitch = irs_ffd_row_pitch(n, sk);
            long long rb = 2048 / n;
            rb = rb < 1 ? 1 : (rb > 16 ? 16 : rb);
            const size_t smem = (size_t)(rb * pitch) * sizeof(float);
            if (smem <= 48 * 1024) {
                long long blocks = (rows + rb - 1) / rb;
                if (blocks > 148 * 8) blocks = 148 * 8;
                ffd_bwd_rows_kernel<<<(unsigned)blocks, 256, smem, st>>>(src, dst, (unsigned)rows, (int)rb, g, n, ax);
                IRS_LAUNCH_CHECK();
                continue;
            }
        }
        // Generic kernel.  Four outputs per thread only when that still fills the GPU (the coarse-grid passes are latency-
        // bound otherwise), and never for an adjoint along the contiguous axis: its lanes read windows s elements apart, so
        // one control point per lane keeps a warp's loads within 32 s floats, four per lane would spread them four times wider.
        const bool vec = total % 4 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && !(adjoint && inner == 1) &&
                         total >= (1ll << 20);
        const unsigned groups = (unsigned)(vec ? total / 4 : total);
        unsigned blocks = (groups + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        if (vec) {
            if (adjoint) ffd_axis_kernel<4, true><<<blocks, 256, 0, st>>>(src, dst, groups, g, n, (unsigned)inner, ax);
            else ffd_axis_kernel<4, false><<<blocks, 256, 0, st>>>(src, dst, groups, g, n, (unsigned)inner, ax);
        } else {
            if (adjoint) ffd_axis_kernel<1, true><<<blocks, 256, 0, st>>>(src, dst, groups, g, n, (unsigned)inner, ax);
            else ffd_axis_kernel<1, false><<<blocks, 256, 0, st>>>(src, dst, groups, g, n, (unsigned)inner, ax);
        }
        IRS_LAUNCH_CHECK();
    }
    return IRS_OK;
}

int check_ffd(int C, int gD, int gH, int gW, int D, int H, int W) {
    IRS_CHECK_DIMS(C, D, H, W);
    if (gD < 1 || gH < 1 || gW < 1) return IRS_ERR_BAD_ARG;
    return IRS_OK;
}

}  // namespace

int irs_launch_ffd(const float* in, float* out, bool adjoint, const float (*kernels)[32], const int* cps, float* work,
                   int C, IrsDims g, IrsDims d, cudaStream_t st) {
    IrsFfdAxis az, ay, ax;
    IRS_TRY(make_axis(az, kernels[0], cps[0], cps[0]));
    IRS_TRY(make_axis(ay, kernels[1], cps[1], cps[1]));
    IRS_TRY(make_axis(ax, kernels[2], cps[2], cps[2]));
    float* t1 = work;                                       // (C 3, D, gH, gW)
    float* t2 = work + (size_t)C * 3 * d.D * g.H * g.W;     // (C 3, D, H, gW)
    if (!adjoint) {
        IRS_TRY(launch_axis(in, t1, false, (long long)C * 3, g.D, d.D, (long long)g.H * g.W, az, st));
        IRS_TRY(launch_axis(t1, t2, false, (long long)C * 3 * d.D, g.H, d.H, g.W, ay, st));
        return launch_axis(t2, out, false, (long long)C * 3 * d.D * d.H, g.W, d.W, 1, ax, st);
    }
    IRS_TRY(launch_axis(in, t2, true, (long long)C * 3 * d.D * d.H, g.W, d.W, 1, ax, st));
    IRS_TRY(launch_axis(t2, t1, true, (long long)C * 3 * d.D, g.H, d.H, g.W, ay, st));
    return launch_axis(t1, out, true, (long long)C * 3, g.D, d.D, (long long)g.H * g.W, az, st);
}

extern "C" size_t irs_ffd_work_floats(int C, int gD, int gH, int gW, int D, int H, int W) {
    if (C < 1 || gD < 1 || gH < 1 || gW < 1 || D < 1 || H < 1 || W < 1) return 0;
    return (size_t)C * 3 * ((size_t)D * gH * gW + (size_t)D * H * gW);
}

extern "C" int irs_bspline_axis(const float* in, float* out, int adjoint, long long outer, int g, int n, long long inner,
                                const float* kernel_host, int stride, int crop_start, void* stream) {
    if (!in || !out) return IRS_ERR_BAD_ARG;
    IrsFfdAxis ax;
    IRS_TRY(make_axis(ax, kernel_host, stride, crop_start));
    return launch_axis(in, out, adjoint != 0, outer, g, n, inner, ax, (cudaStream_t)stream);
}

static int ffd_entry(const float* in, float* out, bool adjoint, const float* kd, const float* kh, const float* kw, int sD,
                     int sH, int sW, float* work, int C, int gD, int gH, int gW, int D, int H, int W, void* stream) {
    IRS_TRY(check_ffd(C, gD, gH, gW, D, H, W));
    if (!in || !out || !work || !kd || !kh || !kw) return IRS_ERR_BAD_ARG;
    const int cps[3] = {sD, sH, sW};
    const float* src[3] = {kd, kh, kw};
    float kernels[3][32];
    for (int a = 0; a < 3; ++a) {
        if (cps[a] < 1 || cps[a] > IRS_FFD_MAX_STRIDE) return IRS_ERR_BAD_ARG;
        for (int j = 0; j < 32; ++j) kernels[a][j] = j < 4 * cps[a] - 1 ? src[a][j] : 0.f;
    }
    return irs_launch_ffd(in, out, adjoint, kernels, cps, work, C, IrsDims{gD, gH, gW}, IrsDims{D, H, W},
                          (cudaStream_t)stream);
}

extern "C" int irs_ffd_fwd(const float* cp, const float* kernel_d_host, const float* kernel_h_host,
                           const float* kernel_w_host, int sD, int sH, int sW, float* work, float* dense, int C, int gD,
                           int gH, int gW, int D, int H, int W, void* stream) {
    return ffd_entry(cp, dense, false, kernel_d_host, kernel_h_host, kernel_w_host, sD, sH, sW, work, C, gD, gH, gW, D, H,
                     W, stream);
}

extern "C" int irs_ffd_bwd(const float* g_dense, const float* kernel_d_host, const float* kernel_h_host,
                           const float* kernel_w_host, int sD, int sH, int sW, float* work, float* g_cp, int C, int gD,
                           int gH, int gW, int D, int H, int W, void* stream) {
    return ffd_entry(g_dense, g_cp, true, kernel_d_host, kernel_h_host, kernel_w_host, sD, sH, sW, work, C, gD, gH, gW, D,
                     H, W, stream);
}
