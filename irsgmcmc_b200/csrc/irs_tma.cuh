// irs_tma.cuh -- TMA (cp.async.bulk.tensor) + mbarrier plumbing for the plane-marching kernels (sm_100a).
//
// A planar vector field (C,3,D,H,W) is described to the TMA unit as the rank-4 tensor {W, H, D, 3C}; a kernel asks for
// the box {BW, BH, 1, 3} = one z-plane of a tile with its halo, all three components, in ONE instruction issued by one
// thread.  Elements outside the volume (negative coordinates, halo beyond the border, planes -1 / D) arrive as zeros,
// so the marching kernels carry no bounds logic for their loads and spend no issue slots on address arithmetic.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>

// ---- host: tensor-map encoding through the runtime's driver entry point (no link-time dependency on libcuda) ----------
inline PFN_cuTensorMapEncodeTiled_v12000 irs_tma_encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    }
    return fn;
}

// true when the field can be described to the TMA unit: 16-byte aligned base and row pitch
inline bool irs_tma_field_ok(const void* base, int W) {
    return (W % 4) == 0 && (reinterpret_cast<uintptr_t>(base) % 16) == 0 && irs_tma_encoder() != nullptr;
}

// tensor map of a planar fp32 field with `nch` (= 3 * chains) channel volumes of D x H x W; box = {bw, bh, 1, 3}
inline int irs_tma_encode_field(CUtensorMap* map, const float* base, int nch, int D, int H, int W, int bw, int bh) {
    PFN_cuTensorMapEncodeTiled_v12000 enc = irs_tma_encoder();
    if (enc == nullptr) return -1;
    const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)nch};
    const cuuint64_t gstride[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * D * 4};
    const cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u, 3u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

// ---- device ------------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t irs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void irs_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(irs_smem_u32(bar)), "r"(count) : "memory");
}
// makes the barrier initialisation visible to the async proxy (the TMA unit) before the first copy is issued
__device__ __forceinline__ void irs_mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void irs_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(irs_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool irs_mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(irs_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void irs_mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!irs_mbar_try_wait(bar, parity)) {}
}

// one z-plane box of a planar field: coordinates (x, y, z, channel) of the box corner, may lie outside the tensor
__device__ __forceinline__ void irs_tma_load_plane(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y,
                                                   int z, int ch) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(irs_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(irs_smem_u32(bar)), "r"(x), "r"(y),
        "r"(z), "r"(ch)
        : "memory");
}
__device__ __forceinline__ void irs_tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// Programmatic dependent launch: `wait` blocks until the preceding kernel of the stream has completed and its writes are
// visible; `launch_dependents` lets the next kernel's CTAs be scheduled as soon as SM resources free up.
__device__ __forceinline__ void irs_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void irs_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
