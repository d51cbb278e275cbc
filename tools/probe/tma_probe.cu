// development probe: does one TMA plane-box load behave as irs_tma.cuh assumes? (not part of the product)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../irsgmcmc_b200/csrc/irs_tma.cuh"

__global__ void probe_kernel(const __grid_constant__ CUtensorMap tmap, float* out, int n_out, int x, int y, int z, int ch,
                             unsigned bytes) {
    extern __shared__ __align__(128) float smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + n_out);
    if (threadIdx.x == 0) { irs_mbar_init(bar, 1); irs_mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { irs_mbar_expect_tx(bar, bytes); irs_tma_load_plane(smem, &tmap, bar, x, y, z, ch); }
    irs_mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < n_out; i += blockDim.x) out[i] = smem[i];
}

int run(int n, int bw, int bh, int x, int y, int z) {
    const int nch = 3;
    std::vector<float> h((size_t)nch * n * n * n);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i + 1);
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    const int n_out = 3 * bh * bw;
    cudaMalloc(&o, n_out * 4);
    CUtensorMap map;
    int r = irs_tma_encode_field(&map, d, nch, n, n, n, bw, bh);
    printf("n=%d box=%dx%d at (%d,%d,%d): encode=%d ", n, bw, bh, x, y, z, r);
    if (r != 0) { printf("\n"); return 1; }
    probe_kernel<<<1, 128, n_out * 4 + 64>>>(map, o, n_out, x, y, z, 0, (unsigned)n_out * 4);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel=%s ", cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 2; }
    std::vector<float> ho(n_out);
    cudaMemcpy(ho.data(), o, n_out * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < 3; ++c)
        for (int j = 0; j < bh; ++j)
            for (int i = 0; i < bw; ++i) {
                const int gx = x + i, gy = y + j;
                float want = 0.f;
                if (gx >= 0 && gx < n && gy >= 0 && gy < n && z >= 0 && z < n) want = h[(((size_t)c * n + z) * n + gy) * n + gx];
                if (ho[(c * bh + j) * bw + i] != want) ++bad;
            }
    printf("mismatches=%d\n", bad);
    cudaFree(d); cudaFree(o);
    return bad != 0;
}

int main(int argc, char** argv) {
    if (argc < 7) return 9;
    return run(atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]));
}
