// irs_bodies.cuh -- per-voxel bodies of the gather kernels as __host__ __device__ functions: the CUDA kernels call them
// once per thread; tests/host_emul.cu calls them in a plain loop on the CPU to check the arithmetic against the oracle
// (test infrastructure only -- libirsgmcmc.so contains no host path).
#pragma once
#include "irs_common.cuh"

#ifdef __CUDA_ARCH__
#define IRS_LDG(p) __ldg(p)
#else
#define IRS_LDG(p) (*(p))
#endif

IRS_HD void irs_voxel_xyz(long long i, IrsDims d, int& x, int& y, int& z) {
    x = (int)(i % d.W); y = (int)((i / d.W) % d.H); z = (int)(i / ((long long)d.W * d.H));
}

// one scaling-and-squaring step for voxel i of one chain; returns max |u(i)|
IRS_HD float irs_body_svf_fwd(const float* __restrict__ u, float in_scale, float* __restrict__ o, long long V,
                              long long i, IrsDims d) {
    int x, y, z;
    irs_voxel_xyz(i, d, x, y, z);
    const float ux = u[i] * in_scale, uy = u[V + i] * in_scale, uz = u[2 * V + i] * in_scale;
    const float px = irs_clampf((float)x + ux, 0.f, (float)(d.W - 1));
    const float py = irs_clampf((float)y + uy, 0.f, (float)(d.H - 1));
    const float pz = irs_clampf((float)z + uz, 0.f, (float)(d.D - 1));
    const IrsCell cell = irs_cell(px, py, pz, d);
    o[i] = ux + in_scale * irs_interp(cell, [&](int k) { return IRS_LDG(u + k); });
    o[V + i] = uy + in_scale * irs_interp(cell, [&](int k) { return IRS_LDG(u + V + k); });
    o[2 * V + i] = uz + in_scale * irs_interp(cell, [&](int k) { return IRS_LDG(u + 2 * V + k); });
    return fmaxf(fabsf(ux), fmaxf(fabsf(uy), fabsf(uz)));
}

// one adjoint step for target voxel i: direct + position term, and (gather_radius >= 0) the interpolation transpose
// gathered over the window of that radius
IRS_HD void irs_body_svf_bwd(const float* __restrict__ u, float in_scale, const float* __restrict__ gp,
                             float* __restrict__ g, int gather_radius, float out_scale, long long V, long long i,
                             IrsDims d) {
    int x, y, z;
    irs_voxel_xyz(i, d, x, y, z);
    const float g0 = gp[i], g1 = gp[V + i], g2 = gp[2 * V + i];
    float ax = g0, ay = g1, az = g2;
    {   // position term
        float px = (float)x + u[i] * in_scale, py = (float)y + u[V + i] * in_scale, pz = (float)z + u[2 * V + i] * in_scale;
        const float mx = irs_inside(px, d.W), my = irs_inside(py, d.H), mz = irs_inside(pz, d.D);
        px = irs_clampf(px, 0.f, (float)(d.W - 1));
        py = irs_clampf(py, 0.f, (float)(d.H - 1));
        pz = irs_clampf(pz, 0.f, (float)(d.D - 1));
        const IrsCell cell = irs_cell(px, py, pz, d);
        float jx = 0.f, jy = 0.f, jz = 0.f, dx, dy, dz;
        irs_interp_grad(cell, [&](int k) { return IRS_LDG(u + k); }, dx, dy, dz);
        jx += g0 * dx; jy += g0 * dy; jz += g0 * dz;
        irs_interp_grad(cell, [&](int k) { return IRS_LDG(u + V + k); }, dx, dy, dz);
        jx += g1 * dx; jy += g1 * dy; jz += g1 * dz;
        irs_interp_grad(cell, [&](int k) { return IRS_LDG(u + 2 * V + k); }, dx, dy, dz);
        jx += g2 * dx; jy += g2 * dy; jz += g2 * dz;
        ax += mx * in_scale * jx; ay += my * in_scale * jy; az += mz * in_scale * jz;
    }
    if (gather_radius >= 0) {  // interpolation transpose as a gather, candidates pruned axis by axis
        const int R = gather_radius;
        const int z_lo = z - R > 0 ? z - R : 0, z_hi = z + R < d.D - 1 ? z + R : d.D - 1;
        const int y_lo = y - R > 0 ? y - R : 0, y_hi = y + R < d.H - 1 ? y + R : d.H - 1;
        const int x_lo = x - R > 0 ? x - R : 0, x_hi = x + R < d.W - 1 ? x + R : d.W - 1;
        const float xmax = (float)(d.W - 1), ymax = (float)(d.H - 1), zmax = (float)(d.D - 1);
        for (int sz = z_lo; sz <= z_hi; ++sz) {
            for (int sy = y_lo; sy <= y_hi; ++sy) {
                const long long row = ((long long)sz * d.H + sy) * d.W;
                for (int sx = x_lo; sx <= x_hi; ++sx) {
                    const long long s = row + sx;
                    const float wx = irs_hat(irs_clampf((float)sx + IRS_LDG(u + s) * in_scale, 0.f, xmax), x);
                    if (wx == 0.f) continue;
                    const float wy = irs_hat(irs_clampf((float)sy + IRS_LDG(u + V + s) * in_scale, 0.f, ymax), y);
                    if (wy == 0.f) continue;
                    const float wz = irs_hat(irs_clampf((float)sz + IRS_LDG(u + 2 * V + s) * in_scale, 0.f, zmax), z);
                    if (wz == 0.f) continue;
                    const float w = wx * wy * wz;
                    ax += w * IRS_LDG(gp + s); ay += w * IRS_LDG(gp + V + s); az += w * IRS_LDG(gp + 2 * V + s);
                }
            }
        }
    }
    g[i] = ax * out_scale; g[V + i] = ay * out_scale; g[2 * V + i] = az * out_scale;
}

// scatter form of the interpolation transpose for source voxel i; `add(index, channel, value)` accumulates
template <typename ADD>
IRS_HD void irs_body_svf_bwd_scatter(const float* __restrict__ u, float in_scale, const float* __restrict__ gp,
                                     float out_scale, long long V, long long i, IrsDims d, ADD add) {
    int x, y, z;
    irs_voxel_xyz(i, d, x, y, z);
    const float px = irs_clampf((float)x + u[i] * in_scale, 0.f, (float)(d.W - 1));
    const float py = irs_clampf((float)y + u[V + i] * in_scale, 0.f, (float)(d.H - 1));
    const float pz = irs_clampf((float)z + u[2 * V + i] * in_scale, 0.f, (float)(d.D - 1));
    const IrsCell cell = irs_cell(px, py, pz, d);
    const float g0 = gp[i] * out_scale, g1 = gp[V + i] * out_scale, g2 = gp[2 * V + i] * out_scale;
    for (int corner = 0; corner < 8; ++corner) {
        const int bx = corner & 1, by = (corner >> 1) & 1, bz = corner >> 2;
        if ((bx && !cell.sx) || (by && !cell.sy) || (bz && !cell.sz)) continue;  // out-of-volume corner: weight 0
        const float w = (bx ? cell.fx : 1.f - cell.fx) * (by ? cell.fy : 1.f - cell.fy) * (bz ? cell.fz : 1.f - cell.fz);
        if (w == 0.f) continue;
        const long long t = cell.i000 + bx * cell.sx + by * cell.sy + bz * cell.sz;
        add(t, 0, w * g0); add(t, 1, w * g1); add(t, 2, w * g2);
    }
}

// position of output voxel i from a normalised grid T, in ATen's operation order (no jitter)
IRS_HD void irs_position_from_T(const float* __restrict__ T, long long V, long long i, IrsDims d, float& px, float& py,
                                float& pz) {
    px = irs_unnormalise(T[i], d.W);
    py = irs_unnormalise(T[V + i], d.H);
    pz = irs_unnormalise(T[2 * V + i], d.D);
}

IRS_HD float irs_body_warp_fwd(const float* __restrict__ im, float px, float py, float pz, IrsDims d) {
    px = irs_clampf(px, 0.f, (float)(d.W - 1));
    py = irs_clampf(py, 0.f, (float)(d.H - 1));
    pz = irs_clampf(pz, 0.f, (float)(d.D - 1));
    const IrsCell cell = irs_cell(px, py, pz, d);
    return irs_interp(cell, [&](int k) { return IRS_LDG(im + k); });
}

// d out / d position (zero on/outside the border), scaled per axis by (sx, sy, sz)
IRS_HD void irs_body_warp_grad(const float* __restrict__ im, float px, float py, float pz, IrsDims d, float go,
                               float sx, float sy, float sz, float& gx, float& gy, float& gz) {
    const float mx = irs_inside(px, d.W) * sx, my = irs_inside(py, d.H) * sy, mz = irs_inside(pz, d.D) * sz;
    px = irs_clampf(px, 0.f, (float)(d.W - 1));
    py = irs_clampf(py, 0.f, (float)(d.H - 1));
    pz = irs_clampf(pz, 0.f, (float)(d.D - 1));
    const IrsCell cell = irs_cell(px, py, pz, d);
    float dx, dy, dz;
    irs_interp_grad(cell, [&](int k) { return IRS_LDG(im + k); }, dx, dy, dz);
    gx = go * dx * mx; gy = go * dy * my; gz = go * dz * mz;
}

IRS_HD long long irs_body_nearest_index(const float* __restrict__ T, long long V, long long i, IrsDims d) {
    const int ix = irs_nearest_coord(T[i], d.W);
    const int iy = irs_nearest_coord(T[V + i], d.H);
    const int iz = irs_nearest_coord(T[2 * V + i], d.D);
    return ((long long)iz * d.H + iy) * d.W + ix;
}
