/* TEST INFRASTRUCTURE -- plain-C restatement of the reference's nearest-neighbour segmentation / mask warp.
 *
 * Reference: RegistrationModule.forward, seg/mask branch (utils/registration.py:20-27 of dgrzech/ir-sgmcmc):
 *   F.grid_sample(seg.float(), T.permute(0,2,3,4,1), mode='nearest', padding_mode='border', align_corners=True)
 * The arithmetic lives in PyTorch (un-pinned by the reference; 2.11.0 here), ATen GridSampler:
 *   unnormalise  x = ((g + 1) / 2) * (size - 1)            (GridSampler.cuh:21-31 / GridSampler.h, align_corners)
 *   clip         x = min(size - 1, max(x, 0))              (:53-57)
 *   round        i = nearbyint(x)   (round half to even)   (nearest branch)
 * all in fp32.  Integer result => the CUDA kernel must match bit for bit.  NaN coordinates are out of contract
 * (ATen's CPU and CUDA paths disagree on them).
 *
 * Compiled with -ffp-contract=off so that no FMA changes the rounding.  Only tests/ load this library.
 */
#include <math.h>
#include <stdint.h>

static int nearest_coord(float g, int n) {
    volatile float a = g + 1.0f;
    volatile float b = a / 2.0f;
    volatile float x = b * (float)(n - 1);
    float lo = x > 0.0f ? x : 0.0f;          /* max(x, 0) */
    float hi = (float)(n - 1);
    float c = lo < hi ? lo : hi;              /* min(size - 1, .) */
    return (int)nearbyintf(c);                /* default rounding mode: half to even */
}

/* T: (C,3,D,H,W) normalised grid, channel 0 = x (W axis); seg: (1 or C, D,H,W) int16; out: (C,D,H,W) */
void oracle_warp_nearest_i16(const int16_t* seg, long long seg_chain_stride, const float* T, int16_t* out, int C, int D,
                             int H, int W) {
    const long long V = (long long)D * H * W;
    for (int c = 0; c < C; ++c) {
        const float* Tc = T + (long long)c * 3 * V;
        for (long long i = 0; i < V; ++i) {
            const int ix = nearest_coord(Tc[i], W), iy = nearest_coord(Tc[V + i], H), iz = nearest_coord(Tc[2 * V + i], D);
            out[c * V + i] = seg[c * seg_chain_stride + ((long long)iz * H + iy) * W + ix];
        }
    }
}

void oracle_warp_nearest_u8(const uint8_t* seg, long long seg_chain_stride, const float* T, uint8_t* out, int C, int D,
                            int H, int W) {
    const long long V = (long long)D * H * W;
    for (int c = 0; c < C; ++c) {
        const float* Tc = T + (long long)c * 3 * V;
        for (long long i = 0; i < V; ++i) {
            const int ix = nearest_coord(Tc[i], W), iy = nearest_coord(Tc[V + i], H), iz = nearest_coord(Tc[2 * V + i], D);
            out[c * V + i] = seg[c * seg_chain_stride + ((long long)iz * H + iy) * W + ix];
        }
    }
}
