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

// position inside an (outer, len, inner) row-major array; flat indices stay below 2^32 elements per launch (checked by the
// launcher), so the two divisions are 32-bit -- the 64-bit ones cost more than the interpolation itself
struct IrsFfdPos {
    unsigned o, x, in_i;
};

IRS_HD IrsFfdPos irs_ffd_decompose(unsigned idx, unsigned len, unsigned inner) {
    IrsFfdPos p;
    const unsigned rest = idx / inner;
    p.in_i = idx - rest * inner;
    p.o = rest / len;
    p.x = rest - p.o * len;
    return p;
}

// the next flat index
IRS_HD void irs_ffd_advance(IrsFfdPos& p, unsigned len, unsigned inner) {
    if (++p.in_i == inner) {
        p.in_i = 0;
        if (++p.x == len) { p.x = 0; ++p.o; }
    }
}

// dense element p.x of row (p.o, p.in_i) from the g control points of that row; p.x + off + 2 s - 1 = q s + r
IRS_HD float irs_body_ffd_axis_fwd(const float* __restrict__ cp, IrsFfdPos p, int q, int r, int g, unsigned inner,
                                   const IrsFfdAxis& ax) {
    const float* base = cp + ((size_t)p.o * g) * inner + p.in_i;
    float acc = 0.f;
#pragma unroll
    for (int m = 3; m >= 0; --m) {   // ascending control point index, the order conv_transpose1d accumulates in
        const int i = q - m, j = r + m * ax.s;
        if (i >= 0 && i < g && j <= 4 * ax.s - 2) acc = fmaf(ax.k[j], base[(size_t)i * inner], acc);
    }
    return acc;
}

// adjoint: control point p.x of row (p.o, p.in_i) sums over the dense elements (n of them) in its support
IRS_HD float irs_body_ffd_axis_bwd(const float* __restrict__ gd, IrsFfdPos p, int n, unsigned inner, const IrsFfdAxis& ax) {
    const int i = (int)p.x;
    // 0 <= t - i s <= 4 s - 2 with t = x + off + 2 s - 1
    int lo = i * ax.s - ax.off - 2 * ax.s + 1, hi = lo + 4 * ax.s - 2;
    lo = lo < 0 ? 0 : lo;
    hi = hi > n - 1 ? n - 1 : hi;
    const float* base = gd + ((size_t)p.o * n) * inner + p.in_i;
    float acc = 0.f;
    for (int x = lo; x <= hi; ++x) {
        const int j = x + ax.off + 2 * ax.s - 1 - i * ax.s;
        acc = fmaf(ax.k[j], base[(size_t)x * inner], acc);
    }
    return acc;
}

// VEC consecutive flat output elements starting at first * VEC (what one CUDA thread computes)
template <int VEC, bool ADJOINT>
IRS_HD void irs_body_ffd_axis_group(const float* __restrict__ in, unsigned first, int g, int n, unsigned inner,
                                    const IrsFfdAxis& ax, float* r) {
    const unsigned len = ADJOINT ? (unsigned)g : (unsigned)n;
    IrsFfdPos p = irs_ffd_decompose(first * VEC, len, inner);
    // (q, rem) of the current dense element follow p.x without further divisions
    const int t = (int)p.x + ax.off + 2 * ax.s - 1;
    int q = t / ax.s, rem = t - q * ax.s;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        r[k] = ADJOINT ? irs_body_ffd_axis_bwd(in, p, n, inner, ax) : irs_body_ffd_axis_fwd(in, p, q, rem, g, inner, ax);
        const unsigned x_before = p.x;
        irs_ffd_advance(p, len, inner);
        if (!ADJOINT && p.x != x_before) {
            if (p.x == 0) {   // next row
                const int t0 = ax.off + 2 * ax.s - 1;
                q = t0 / ax.s;
                rem = t0 - q * ax.s;
            } else if (++rem == ax.s) {
                rem = 0;
                ++q;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Table forms used by the fast kernels.  Everything that depends only on the coordinate along the axis is computed once
// per block: per dense element the four control-point indices (ascending, clamped into the grid) and their weights
// (zero where the generic body skips a term), so an output costs two table reads, four loads and four FMAs.
// ---------------------------------------------------------------------------------------------------------------------
struct IrsFfdEntry {
    int i[4];     // q - 3 .. q, clamped to [0, g - 1]
    float w[4];   // taps[r + 3 s], taps[r + 2 s], taps[r + s], taps[r]; 0 for a term outside the grid / the kernel
};

IRS_HD IrsFfdEntry irs_ffd_entry(int x, int g, const IrsFfdAxis& ax) {
    const int t = x + ax.off + 2 * ax.s - 1;
    const int q = t / ax.s, r = t - q * ax.s;
    IrsFfdEntry e;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int m = 3 - a, i = q - m, j = r + m * ax.s;
        const bool ok = i >= 0 && i < g && j <= 4 * ax.s - 2;
        e.i[a] = i < 0 ? 0 : (i > g - 1 ? g - 1 : i);
        e.w[a] = ok ? ax.k[j] : 0.f;
    }
    return e;
}

// same accumulation order as irs_body_ffd_axis_fwd (ascending control point index)
IRS_HD float irs_body_ffd_axis_fwd_tab(const float* __restrict__ row, unsigned inner, const IrsFfdEntry& e) {
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) acc = fmaf(e.w[a], row[(size_t)e.i[a] * inner], acc);
    return acc;
}

// Adjoint along the contiguous axis from a staged copy of the dense row: element x of the row sits at irs_ffd_skew(x).
// With an even spacing the lanes of a warp (consecutive control points) would read shared-memory words s apart; the skew
// x + (x >> tz), tz = trailing zero bits of s, turns that into s + s / 2^tz, which is odd: no bank conflicts.
struct IrsFfdSkew {
    int tz, mask;   // mask = 0 switches the skew off (odd spacing)
};

IRS_HD IrsFfdSkew irs_ffd_make_skew(int s) {
    IrsFfdSkew k;
    k.tz = 0;
    k.mask = (s & 1) ? 0 : -1;
    if (k.mask) while (!((s >> k.tz) & 1)) ++k.tz;
    return k;
}
IRS_HD int irs_ffd_skew(int x, IrsFfdSkew k) { return x + ((x >> k.tz) & k.mask); }
IRS_HD int irs_ffd_row_pitch(int n, IrsFfdSkew k) { return irs_ffd_skew(n - 1, k) + 1; }

// same accumulation order as irs_body_ffd_axis_bwd (ascending dense element)
IRS_HD float irs_body_ffd_axis_bwd_row(const float* __restrict__ srow, IrsFfdSkew k, int i, int n, const IrsFfdAxis& ax) {
    int lo = i * ax.s - ax.off - 2 * ax.s + 1, hi = lo + 4 * ax.s - 2;
    int j = lo < 0 ? -lo : 0;   // first tap that falls on the row
    lo = lo < 0 ? 0 : lo;
    hi = hi > n - 1 ? n - 1 : hi;
    float acc = 0.f;
    for (int x = lo; x <= hi; ++x, ++j) acc = fmaf(ax.k[j], srow[irs_ffd_skew(x, k)], acc);
    return acc;
}
