// common.cuh -- shared helpers for the sm_100a kernels of libhp_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "hp_b200.h"
#ifdef HP_BENCH_BUILD
#include "hp_b200_bench.h"
#endif

namespace hp {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);  // returns HP_OK or HP_ERR_CUDA (and records the message)
int sm_count();                                     // multiprocessor count of the current device (cached per device)

// Opt a kernel in to > 48 KB of dynamic shared memory once per device.
struct SmemAttrCache {
    size_t bytes[64] = {0};
};
template <typename K>
cudaError_t ensure_dynamic_smem(K kern, size_t bytes, SmemAttrCache &cache) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && cache.bytes[dev] >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess && dev >= 0 && dev < 64) cache.bytes[dev] = bytes;
    return e;
}

#define HP_REQUIRE(cond, ...)               \
    do {                                    \
        if (!(cond)) {                      \
            hp::set_error(__VA_ARGS__);     \
            return HP_ERR_INVALID_ARGUMENT; \
        }                                   \
    } while (0)

#define HP_CUDA(call)                                    \
    do {                                                 \
        int _hp_rc = hp::check_cuda((call), #call);      \
        if (_hp_rc != HP_OK) return _hp_rc;              \
    } while (0)

#define HP_LAUNCH_CHECK(name)                                        \
    do {                                                             \
        int _hp_rc = hp::check_cuda(cudaGetLastError(), name);       \
        if (_hp_rc != HP_OK) return _hp_rc;                          \
    } while (0)

// ---- packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2) -----------------------------
// Each lane of the pair is an ordinary IEEE round-to-nearest fp32 op, so results are
// bit-identical to the scalar __fsub_rn / __fmul_rn / __fmaf_rn sequence.
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// 3-input minimum (sm_100 FMNMX3).  NaN operands are ignored like fminf.
__device__ __forceinline__ float min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// The one squared-distance association of the whole library (== reference SASS):
//   d = fma(dz, dz, fma(dx, dx, dy*dy)),   dx = cand - query
__device__ __forceinline__ float sqdist_exact(float qx, float qy, float qz, float cx, float cy, float cz) {
    float dx = __fsub_rn(cx, qx), dy = __fsub_rn(cy, qy), dz = __fsub_rn(cz, qz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}
__device__ __forceinline__ f32x2 sqdist_exact2(f32x2 qx, f32x2 qy, f32x2 qz, f32x2 cx, f32x2 cy, f32x2 cz) {
    f32x2 dx = sub2(cx, qx), dy = sub2(cy, qy), dz = sub2(cz, qz);
    return fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
}

// ---- mbarrier + 1-D bulk (TMA) global->shared copy ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "HP_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra HP_DONE_%=;\n\t"
        "bra HP_WAIT_%=;\n\t"
        "HP_DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// dst (shared), src (global) 16-byte aligned, bytes a multiple of 16.  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// order prior generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace hp
