// irs_ffd_body.cuh -- cubic B-spline free-form deformation along ONE axis, as __host__ __device__ bodies (the CUDA kernels
// of irs_ffd.cu call them once per thread; tests/host_emul.cu calls them in plain loops on the CPU).
//
// Reference: Cubic_B_spline_FFD_3D.forward, utils/transformation.py:132-152 = per axis a transposed 1-D convolution
// (F.conv_transpose1d, stride s, kernel K of 4 s - 1 taps from B_spline_1D_kernel :95-103, padding 2 s - 1) followed by
// the crop [s, s + n).  Written out, element p of the un-cropped result is
//     full[p] = sum_i cp[i] * K[p + 2 s - 1 - i s],        0 <= p + 2 s - 1 - i s <= 4 s - 2
// so with t = p + 2 s - 1 = q s + r (0 <= r < s) the contributing control points are i = q - m, m = 0..3, with weight
// K[r + m s] (absent when r + m s > 4 s - 2, i.e. m = 3 and r = s - 1).  `off` is the crop start (s in the module, 0 for
// the un-cropped conv1D), the output element x is p = x + off.
#pragma once
#include "irs_common.cuh"

#define IRS_FFD_MAX_STRIDE 8
#define IRS_FFD_MAX_KERNEL (4 * IRS_FFD_MAX_STRIDE - 1)

struct IrsFfdAxis {
    int s;                             // control point spacing (stride)
    int off;                           // crop start
    float k[IRS_FFD_MAX_KERNEL + 1];   // the reference's B_spline_1D_kernel(s): 4 s - 1 taps
};

// arrays are (outer, len, inner) row-major; idx enumerates the OUTPUT (outer, n, inner)
IRS_HD float irs_body_ffd_axis_fwd(const float* __restrict__ cp, long long idx, int g, int n, long long inner,
                                   const IrsFfdAxis& ax) {
    const long long in_i = idx % inner;
    const long long rest = idx / inner;
    const int x = (int)(rest % n);
    const long long o = rest / n;
    const int t = x + ax.off + 2 * ax.s - 1;
    const int q = t / ax.s, r = t - q * ax.s;
    const float* base = cp + (o * g) * inner + in_i;
    float acc = 0.f;
#pragma unroll
    for (int m = 3; m >= 0; --m) {   // ascending control point index, the order conv_transpose1d accumulates in
        const int i = q - m, j = r + m * ax.s;
        if (i >= 0 && i < g && j <= 4 * ax.s - 2) acc = fmaf(ax.k[j], base[(long long)i * inner], acc);
    }
    return acc;
}

// adjoint: idx enumerates the control-point side (outer, g, inner); sums over the dense elements in the support
IRS_HD float irs_body_ffd_axis_bwd(const float* __restrict__ gd, long long idx, int g, int n, long long inner,
                                   const IrsFfdAxis& ax) {
    const long long in_i = idx % inner;
    const long long rest = idx / inner;
    const int i = (int)(rest % g);
    const long long o = rest / g;
    // 0 <= t - i s <= 4 s - 2 with t = x + off + 2 s - 1
    int lo = i * ax.s - ax.off - 2 * ax.s + 1, hi = lo + 4 * ax.s - 2;
    lo = lo < 0 ? 0 : lo;
    hi = hi > n - 1 ? n - 1 : hi;
    const float* base = gd + (o * n) * inner + in_i;
    float acc = 0.f;
    for (int x = lo; x <= hi; ++x) {
        const int j = x + ax.off + 2 * ax.s - 1 - i * ax.s;
        acc = fmaf(ax.k[j], base[(long long)x * inner], acc);
    }
    return acc;
}
