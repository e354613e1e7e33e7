// target_network_tc5.cuh -- the TargetNetwork FORWARD (3 -> 32 -> 64 -> 128 -> 64 -> 3) on the 5th-generation tensor cores:
// tcgen05.mma kind::tf32, accumulators AND activations in tensor memory (included by target_network.cu; opt-in, mode 2).
//
// Same arithmetic as target_network_mma.cuh -- error-compensated 3xTF32, every fp32 product as lo*hi + hi*lo + hi*hi with fp32
// accumulation -- on the tcgen05 path, which runs tf32 at four times the rate of legacy mma.sync.  The point of the design is that
// NO activation ever touches shared memory or registers of another thread:
//   * a 128-point tile is the M dimension: point p of the tile is lane p of tensor memory.  The accumulator of layer L
//     (D[128 x N] fp32, N columns) is read by the thread that owns the lane (tcgen05.ld 32x32b: thread = point, all channels),
//     gets bias + ReLU, is split into hi (the value itself: the tensor core truncates) and lo = x - trunc_tf32(x), and goes back to
//     tensor memory (tcgen05.st) as the A OPERAND of layer L+1 -- tcgen05.mma takes A straight from tensor memory;
//   * the B operands are the sample's weight matrices, pre-split into hi / lo once per sample and laid out in shared memory in the
//     canonical K-major no-swizzle form (8 x 16-byte core matrices; LBO = next 16-byte K chunk, SBO = next 8 rows);
//   * layer 1 (K = 3) and layer 5 (N = 3) are plain FFMA in the owning thread.
// One CTA per SM: two 256-column tile slots (all 512 columns of tensor memory), each served by four epilogue warps; a ninth warp
// issues the MMAs of both slots in turn, so one slot's epilogue runs under the other slot's MMAs.  Hand-over is by mbarriers
// (epilogue -> MMA: 128 arrivals after tcgen05.wait::st; MMA -> epilogue: tcgen05.commit).
// Column plan of a slot (R0 = columns 0-127, R1 = 128-255):
//   A1 hi|lo R1[0,64) -> D2 R0[0,64) -> A2 hi|lo R1[0,128) -> D3 R0[0,128) -> A3a hi R0[0,64) (in place), lo R1[0,64)
//   -> D4 R1[64,128) (K = 0..63) ; A3b hi R0[64,128) (in place), lo R1[0,64) once the first half is consumed -> D4 += (K = 64..127).
#pragma once

namespace hp {

constexpr int T5_EPI_WARPS = 8;                      // 2 slots x 4 warps
constexpr int T5_THREADS = (T5_EPI_WARPS + 1) * 32;  // + the MMA warp
constexpr int T5_W2H = 0;                            // canonical K-major [64][32]
constexpr int T5_W2L = T5_W2H + C2 * C1;
constexpr int T5_W3H = T5_W2L + C2 * C1;             // [128][64]
constexpr int T5_W3L = T5_W3H + C3 * C2;
constexpr int T5_W4H = T5_W3L + C3 * C2;             // [64][128]
constexpr int T5_W4L = T5_W4H + C4 * C3;
constexpr int T5_W1P = T5_W4L + C4 * C3;             // [32][4] = (w0, w1, w2, bias)
constexpr int T5_W5 = T5_W1P + C1 * 4;               // [3][64]
constexpr int T5_B = T5_W5 + 3 * C4;                 // b2[64] b3[128] b4[64] b5[4]
constexpr int T5_FLOATS = T5_B + C2 + C3 + C4 + 4;
constexpr size_t T5_SMEM = (size_t)T5_FLOATS * sizeof(float);
static_assert(T5_SMEM + 1024 <= 227 * 1024 && 2 * T5_SMEM > 227 * 1024, "exactly one CTA per SM (it takes all of tensor memory)");

// canonical K-major no-swizzle position (in floats) of element (n, k) of a [N][KTOT] matrix: 8-row x 16-byte core matrices,
// consecutive K chunks adjacent (LBO = 128 B), 8-row groups KTOT/4 core matrices apart (SBO = KTOT * 32 B)
template <int KTOT>
__device__ __forceinline__ int t5_canon(int n, int k) {
    return ((n >> 3) * (KTOT >> 2) + (k >> 2)) * 32 + (n & 7) * 4 + (k & 3);
}
// W[N][KTOT] (global, row-major) -> hi / lo in canonical order.  Threads walk the CANONICAL positions, so a warp's 32 stores fill one
// 128-byte core matrix (conflict-free) and its 32 loads are eight 16-byte row pieces; 32 loads in flight per thread (L2 latency).
template <int KTOT, int N>
__device__ __forceinline__ void t5_stage_matrix(const float *__restrict__ Wg, float *__restrict__ hi, float *__restrict__ lo, int tid) {
    constexpr int BATCH = 32, TOTAL = N * KTOT;
    for (int c0 = tid; c0 < TOTAL; c0 += BATCH * T5_THREADS) {
        float w[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
            const int c = c0 + u * T5_THREADS;  // canonical position: ((group * KTOT/4 + chunk) * 8 + row) * 4 + kk
            const int kk = c & 3, r = (c >> 2) & 7, cm = c >> 5, chunk = cm % (KTOT / 4), group = cm / (KTOT / 4);
            w[u] = c < TOTAL ? __ldg(Wg + (group * 8 + r) * KTOT + chunk * 4 + kk) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
            const int c = c0 + u * T5_THREADS;
            if (c < TOTAL) {
                hi[c] = w[u];
                lo[c] = w[u] - __uint_as_float(__float_as_uint(w[u]) & 0xffffe000u);
            }
        }
    }
}
// shared-memory matrix descriptor of the K = 8 slice `ks` of a canonical [N][KTOT] matrix
template <int KTOT>
__device__ __forceinline__ uint64_t t5_bdesc(const float *W, int ks) {
    const uint32_t addr = smem_u32(W) + ks * 256;  // two 128-byte core matrices per K = 8 step
    uint64_t d = (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)(128 >> 4) << 16;               // leading byte offset: the second 16-byte K chunk
    d |= (uint64_t)((KTOT * 32) >> 4) << 32;       // stride byte offset: the next 8 rows
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    return d;                                      // base offset 0, layout type 0 (no swizzle)
}
// instruction descriptor: D fp32, A / B tf32, both K-major, M = 128
__device__ __forceinline__ constexpr uint32_t t5_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void t5_mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// one lane of a converged warp (elect.sync): the issuing thread of tcgen05.mma / tcgen05.commit.  The whole warp walks the issue loop
// with warp-uniform operands, so the descriptors stay in uniform registers instead of being moved there per instruction.
__device__ __forceinline__ bool t5_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void t5_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void t5_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[128 x N] (+)= A[128 x 8*NKS] * W^T as three tf32 products; A hi / lo in tensor memory, W hi / lo canonical in shared memory
template <int KTOT, int N>
__device__ __forceinline__ void t5_layer(uint32_t d, uint32_t a_hi, uint32_t a_lo, const float *Wh, const float *Wl, int ks0, int nks,
                                         bool accumulate) {
    constexpr uint32_t idesc = t5_idesc(N);
#pragma unroll 4
    for (int i = 0; i < nks; ++i) {
        const uint64_t bh = t5_bdesc<KTOT>(Wh, ks0 + i), bl = t5_bdesc<KTOT>(Wl, ks0 + i);
        t5_mma(d, a_lo + 8 * i, bh, idesc, (accumulate || i > 0) ? 1u : 0u);
        t5_mma(d, a_hi + 8 * i, bl, idesc, 1u);
        t5_mma(d, a_hi + 8 * i, bh, idesc, 1u);
    }
}
// tensor-memory <-> registers without a wait per instruction: N columns of this thread's lane, one wait for all of them
template <int N>
__device__ __forceinline__ void tmem_load_nowait(uint32_t taddr, uint32_t (&v)[N]) {
#pragma unroll
    for (int o = 0; o < N; o += 16)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : HP_R16(v, o)
                     : "r"(taddr + o));
}
template <int N>
__device__ __forceinline__ void tmem_store_nowait(uint32_t taddr, const uint32_t (&v)[N]) {
#pragma unroll
    for (int o = 0; o < N; o += 16)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr + o),
                     HP_I16(v, o)
                     : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// NC columns of an accumulator -> bias + ReLU -> hi / lo A operands of the next layer (hi may go back in place)
template <int NC>
__device__ __forceinline__ void t5_epilogue(uint32_t src, uint32_t dst_hi, uint32_t dst_lo, const float *__restrict__ bias) {
    uint32_t v[NC], lo[NC];
    tmem_load_nowait<NC>(src, v);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < NC; i += 4) {
        const float4 b4 = *reinterpret_cast<const float4 *>(bias + i);
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float x = fmaxf(__uint_as_float(v[i + u]) + bb[u], 0.f);
            v[i + u] = __float_as_uint(x);
            lo[i + u] = __float_as_uint(x - __uint_as_float(v[i + u] & 0xffffe000u));
        }
    }
    tmem_store_nowait<NC>(dst_hi, v);
    tmem_store_nowait<NC>(dst_lo, lo);
    tmem_wait_st();
}

#ifdef HP_TM_TRACE
#define T5_TRACE(role, tile, ev) do { if (blockIdx.x == 0 && lane == 0 && (tile) < 64) { unsigned long long tns_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns_)); g_tm_trace[role][tile][ev] = tns_; } } while (0)
#else
#define T5_TRACE(role, tile, ev) do { } while (0)
#endif

__global__ void __launch_bounds__(T5_THREADS, 1) tn_tc5_forward_kernel(const TNArgs a) {
    extern __shared__ __align__(16) float sm[];  // no-swizzle descriptors need 16-byte alignment only
    __shared__ uint64_t ready[2], done[2];  // per slot: epilogue -> MMA (128 arrivals), MMA -> epilogue (one commit)
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool mma_warp = warp == T5_EPI_WARPS;
    const int slot = warp >> 2;  // epilogue warps: 0 or 1
    const int ntiles = (a.N + TN_T - 1) / TN_T;
    const long long TT = (long long)a.B * ntiles;
    const long long f0 = (long long)blockIdx.x * TT / gridDim.x, f1 = (long long)(blockIdx.x + 1) * TT / gridDim.x;
    TM_CTA(0);
    if (tid == 0) {
        mbar_init(&ready[0], 128), mbar_init(&ready[1], 128);
        mbar_init(&done[0], 1), mbar_init(&done[1], 1);
    }
    if (mma_warp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    float *w1p = sm + T5_W1P, *w5 = sm + T5_W5, *bs = sm + T5_B;
    uint32_t ready_phase[2] = {0, 0}, done_phase = 0;  // MMA warp: per slot; epilogue warps: their slot's `done`

    long long f = f0;
    while (f < f1) {
        // ---- one sample segment: stage its weights (hi / lo, canonical), then walk its tiles two at a time ----
        const int b = (int)(f / ntiles);
        const long long fe = f1 < (long long)(b + 1) * ntiles ? f1 : (long long)(b + 1) * ntiles;
        const float *wg = a.weights + (size_t)b * a.W;
        __syncthreads();  // everybody is done with the previous sample's weights
        t5_stage_matrix<C1, C2>(wg + a.offw[1], sm + T5_W2H, sm + T5_W2L, tid);
        t5_stage_matrix<C2, C3>(wg + a.offw[2], sm + T5_W3H, sm + T5_W3L, tid);
        t5_stage_matrix<C3, C4>(wg + a.offw[3], sm + T5_W4H, sm + T5_W4L, tid);
        for (int i = tid; i < C1 * 4; i += T5_THREADS) {
            const int o = i >> 2, c = i & 3;
            w1p[i] = c < 3 ? __ldg(wg + a.offw[0] + o * 3 + c) : (a.offb[0] >= 0 ? __ldg(wg + a.offb[0] + o) : 0.f);
        }
        for (int i = tid; i < 3 * C4; i += T5_THREADS) w5[i] = __ldg(wg + a.offw[4] + i);
        for (int i = tid; i < C2 + C3 + C4 + 4; i += T5_THREADS) {
            int l, o;
            if (i < C2) l = 1, o = i;
            else if (i < C2 + C3) l = 2, o = i - C2;
            else if (i < C2 + C3 + C4) l = 3, o = i - C2 - C3;
            else l = 4, o = i - C2 - C3 - C4;
            bs[i] = (a.offb[l] >= 0 && !(l == 4 && o >= 3)) ? __ldg(wg + a.offb[l] + o) : 0.f;
        }
        fence_proxy_async();  // generic-proxy stores above -> async-proxy (tensor core) reads
        __syncthreads();
        if (f == f0) TM_CTA(1);

        if (mma_warp) {
            for (long long t0 = f; t0 < fe; t0 += 2) {
                const int live = (t0 + 1 < fe) ? 2 : 1;
                for (int step = 0; step < 4; ++step) {
                    for (int s = 0; s < live; ++s) {
                        T5_TRACE(1, (int)(t0 - f0) + s, step * 3);
                        mbar_wait(&ready[s], ready_phase[s]);
                        ready_phase[s] ^= 1;
                        T5_TRACE(1, (int)(t0 - f0) + s, step * 3 + 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t R0 = tmem + 256 * s, R1 = R0 + 128;
                        if (t5_elect_one()) {
                            if (step == 0) t5_layer<C1, C2>(R0, R1, R1 + 32, sm + T5_W2H, sm + T5_W2L, 0, C1 / 8, false);
                            else if (step == 1) t5_layer<C2, C3>(R0, R1, R1 + 64, sm + T5_W3H, sm + T5_W3L, 0, C2 / 8, false);
                            else if (step == 2) t5_layer<C3, C4>(R1 + 64, R0, R1, sm + T5_W4H, sm + T5_W4L, 0, 8, false);
                            else t5_layer<C3, C4>(R1 + 64, R0 + 64, R1, sm + T5_W4H, sm + T5_W4L, 8, 8, true);
                            t5_commit(&done[s]);
                        }
                        __syncwarp();
                        T5_TRACE(1, (int)(t0 - f0) + s, step * 3 + 2);
                    }
                }
            }
        } else {
            const uint32_t lanes = (uint32_t)(32 * (warp & 3)) << 16;
            const uint32_t R0 = tmem + lanes + 256 * slot, R1 = R0 + 128;
            const int p_in_tile = 32 * (warp & 3) + lane;
            for (long long t0 = f + slot; t0 < fe; t0 += 2) {
                const int tl = (int)(t0 - (long long)b * ntiles);
                const int p = tl * TN_T + p_in_tile;
                if (warp == 0) T5_TRACE(0, (int)(t0 - f0), 0);
                // ---- layer 1 on the FP32 pipe -> A1 hi | lo ----
                float x0 = 0.f, x1 = 0.f, x2 = 0.f;
                if (p < a.N) {
                    const float *px = a.points + (size_t)b * a.pstride + (size_t)p * 3;
                    x0 = __ldg(px), x1 = __ldg(px + 1), x2 = __ldg(px + 2);
                }
                {
                    uint32_t hi[C1], lo[C1];
#pragma unroll
                    for (int i = 0; i < C1; ++i) {
                        const float4 w = *reinterpret_cast<const float4 *>(w1p + i * 4);
                        const float v = fmaxf(__fmaf_rn(w.z, x2, __fmaf_rn(w.y, x1, __fmaf_rn(w.x, x0, w.w))), 0.f);
                        hi[i] = __float_as_uint(v);
                        lo[i] = __float_as_uint(v - __uint_as_float(hi[i] & 0xffffe000u));
                    }
                    tmem_store_nowait<C1>(R1, hi);
                    tmem_store_nowait<C1>(R1 + 32, lo);
                    tmem_wait_st();
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                t5_mbar_arrive(&ready[slot]);
                if (warp == 0) T5_TRACE(0, (int)(t0 - f0), 1);
                // ---- layer 2 accumulator -> A2 hi | lo ----
                mbar_wait(&done[slot], done_phase), done_phase ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp == 0) T5_TRACE(0, (int)(t0 - f0), 2);
                t5_epilogue<C2>(R0, R1, R1 + 64, bs);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                t5_mbar_arrive(&ready[slot]);
                if (warp == 0) T5_TRACE(0, (int)(t0 - f0), 3);
                // ---- layer 3 accumulator, channels 0-63 -> A3a (hi in place, lo over the dead A2) ----
                mbar_wait(&done[slot], done_phase), done_phase ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp == 0) T5_TRACE(0, (int)(t0 - f0), 4);
                t5_epilogue<64>(R0, R0, R1, bs + C2);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                t5_mbar_arrive(&ready[slot]);
                if (warp == 0) T5_TRACE(0, (int)(t0 - f0), 5);
                // ---- channels 64-127 -> A3b: hi in place now, lo once the MMAs over A3a have read R1[0,64) ----
                {
                    uint32_t v[64], lo[64];
                    tmem_load_nowait<64>(R0 + 64, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 64; ++i) {
                        const float x = fmaxf(__uint_as_float(v[i]) + bs[C2 + 64 + i], 0.f);
                        v[i] = __float_as_uint(x);
                        lo[i] = __float_as_uint(x - __uint_as_float(v[i] & 0xffffe000u));
                    }
                    tmem_store_nowait<64>(R0 + 64, v);
                    mbar_wait(&done[slot], done_phase), done_phase ^= 1;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    tmem_store_nowait<64>(R1, lo);
                    tmem_wait_st();
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                t5_mbar_arrive(&ready[slot]);
                if (warp == 0) T5_TRACE(0, (int)(t0 - f0), 6);
                // ---- layer 4 accumulator -> bias + ReLU -> layer 5 on the FP32 pipe -> out ----
                mbar_wait(&done[slot], done_phase), done_phase ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp == 0) T5_TRACE(0, (int)(t0 - f0), 7);
                float y0 = bs[C2 + C3 + C4], y1 = bs[C2 + C3 + C4 + 1], y2 = bs[C2 + C3 + C4 + 2];
                {
                    uint32_t v[C4];
                    tmem_load_nowait<C4>(R1 + 64, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < C4; ++i) {
                        const float x = fmaxf(__uint_as_float(v[i]) + bs[C2 + C3 + i], 0.f);
                        y0 = __fmaf_rn(x, w5[i], y0), y1 = __fmaf_rn(x, w5[C4 + i], y1), y2 = __fmaf_rn(x, w5[2 * C4 + i], y2);
                    }
                }
                if (warp == 0) T5_TRACE(0, (int)(t0 - f0), 8);
                if (p < a.N) {
                    if (a.channels_first) {
                        float *dst = a.out + (size_t)b * 3 * a.N + p;
                        dst[0] = y0, dst[a.N] = y1, dst[2 * (size_t)a.N] = y2;
                    } else {
                        float *dst = a.out + ((size_t)b * a.N + p) * 3;
                        dst[0] = y0, dst[1] = y1, dst[2] = y2;
                    }
                }
            }
        }
        f = fe;
    }
    TM_CTA(2);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    TM_CTA(3);
    if (mma_warp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace hp
