// target_network_mma.cuh -- the TargetNetwork fast path (3 -> 32 -> 64 -> 128 -> 64 -> 3) on the tensor cores as error-compensated
// 3xTF32 (included by target_network.cu; shares TNArgs and the launch geometry with the FP32-pipe kernels there).
//
// Why it is admissible: tools/tf32x3_study.cu / profiles/r02_tf32x3_study.txt -- with every operand split into hi + lo (both tf32)
// and lo*hi + hi*lo + hi*hi accumulated in fp32, the K = 32..128 contractions of this network stay within 2.3e-6 of the fp32 FFMA
// chain end to end (bar: 1e-5 against torch.mm fp32, model/target_network.py:31-38); a weight-gradient contraction over all 2048 points
// in ONE tensor-core accumulator does not (1.6e-5), so dW is accumulated per 128-point tile on the tensor cores and the tile sums are
// added on the FP32 pipe.
//
// Design (mma.sync.m16n8k8 tf32; the fragment algebra is what makes it cheap):
//   * a WARP owns 16 points and walks ALL layers for them in registers: the accumulator fragment of one layer IS the A fragment of
//     the next.  With MMA column c of channel tile j bound to channel 8j + s(c), s(2t) = t, s(2t+1) = t+4, lane (g,t) holds
//     C = {(g, 8j+t), (g, 8j+t+4), (g+8, 8j+t), (g+8, 8j+t+4)} -- exactly {a0, a2, a1, a3} of contraction step j of the next layer.
//     No activation touches shared memory in the forward.
//   * weights live in shared memory as rows of the NON-contracted index with stride = contraction length + 4 (== 4 mod 32): the B
//     fragment (row 8j + s(g), columns 8k + t and 8k + t + 4) is two conflict-free LDS.32; hi/lo are split on the fly (the FP32 pipe is
//     idle under the MMAs).  The backward stages each matrix a second time TRANSPOSED for the dgrad contraction (over `out`).
//   * layer 1 (K = 3) and the 3-wide contractions of the output layer's backward run as plain FFMA in fragment layout.
//   * BACKWARD (one CTA of 8 warps per SM, 128-point tiles): recompute the forward in registers, storing the activations of the tile to
//     shared memory as [point][channel] rows (stride 296 == 8 mod 32) only because wgrad contracts over POINTS, which needs them
//     transposed: dW_L[o][k] = sum_p Z_L[p][o] A_{L-1}[p][k] reads both operands straight from those rows as conflict-free fragments.
//     The dgrad chain Z_{L-1} = relu'(A_{L-1}) * (Z_L W_L) stays in registers (ReLU gates are 144 bits per lane), Z_L overwrites A_L
//     in shared memory for wgrad.  dW lives in 72 persistent registers per thread across the tiles of a sample; per-CTA partials of a
//     sample shared by several CTAs are folded by the last CTA to arrive in ascending CTA order (deterministic, no float atomics).
#pragma once

namespace hp {

#ifndef HP_TMF_WARPS
#define HP_TMF_WARPS 12
#endif
#ifndef HP_TM_JG
#define HP_TM_JG 4
#endif
constexpr int TMF_WARPS = HP_TMF_WARPS;                 // forward CTA: 384 threads (<= 168 registers each)
constexpr int TMF_THREADS = TMF_WARPS * 32;

// Operand split for 3xTF32.  The tensor core reads only the top 19 bits of an operand register (it TRUNCATES fp32 to tf32), so
//   hi: the fp32 value itself is passed -- the hardware sees trunc_tf32(x);
//   lo: x - trunc_tf32(x), exact in fp32 (one LOP3 + one FADD), of which the hardware again keeps the top 11 significant bits.
// x*y ~ hi_x*hi_y + hi_x*lo_y + lo_x*hi_y with a relative defect <= 2^-20 (mean 2^-22 per operand, towards zero) -- the same order
// as the dropped lo*lo term.  sm_100 has no cvt.rna.tf32 instruction (the PTX cvt expands to four half-rate ALU operations), and
// a round-to-nearest hi would cost a third operation per element for an error budget that is not needed (measured worst case of the
// whole network against float64, tools/tn_error_margins.py: see DESIGN.md 4.3; bar 1e-5).
#ifndef HP_TM_ROUNDED_SPLIT
__device__ __forceinline__ void tf32_split(float x, uint32_t &hi, uint32_t &lo) {
    hi = __float_as_uint(x);
    lo = __float_as_uint(x - __uint_as_float(hi & 0xffffe000u));
}
#else  // hi rounded to nearest (ties away): the truncated bits of x * (1 + 2^-12); FFMA + LOP3 + FADD
__device__ __forceinline__ void tf32_split(float x, uint32_t &hi, uint32_t &lo) {
    hi = __float_as_uint(__fmaf_rn(x, 0x1p-12f, x)) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
#endif
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async4(float *dst_smem, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// W[O][K] (flat, row-major) -> rows of `out` with stride K+4
template <int K, int O, int NT>
__device__ __forceinline__ void tm_stage_nat(const float *__restrict__ Wg, float *__restrict__ dst, int tid) {
    for (int i = tid; i < O * K; i += NT) {
        const int o = i / K, k = i - o * K;
        cp_async4(dst + o * (K + 4) + k, Wg + i);
    }
}
// W[O][K] -> its transpose, rows of `in` with stride O+4
template <int K, int O, int NT>
__device__ __forceinline__ void tm_stage_tr(const float *__restrict__ Wg, float *__restrict__ dst, int tid) {
    for (int i = tid; i < O * K; i += NT) {
        const int o = i / K, k = i - o * K;
        cp_async4(dst + k * (O + 4) + o, Wg + i);
    }
}

// One contraction of the register chain: acc[j] += in (16 points x K, fragment layout) * Ws^T, Ws = rows of the N non-contracted
// channels, stride K+4.  wl = Ws + s(g)*(K+4) + t.  Channel tiles go in groups of JG: the three products of a group are issued
// term by term (lo*hi of all tiles, hi*lo of all, hi*hi of all), so that back-to-back MMAs never share an accumulator.
template <int K, int N>
__device__ __forceinline__ void tm_layer(const float (&in)[K / 8][4], float (&acc)[N / 8][4], const float *__restrict__ wl) {
    constexpr int LD = K + 4;
    constexpr int NTL = N / 8, JG = NTL >= HP_TM_JG ? HP_TM_JG : NTL;
#pragma unroll
    for (int ks = 0; ks < K / 8; ++ks) {
        uint32_t ah[4], al[4];
        tf32_split(in[ks][0], ah[0], al[0]);  // a0 = (g,   t)
        tf32_split(in[ks][2], ah[1], al[1]);  // a1 = (g+8, t)
        tf32_split(in[ks][1], ah[2], al[2]);  // a2 = (g,   t+4)
        tf32_split(in[ks][3], ah[3], al[3]);  // a3 = (g+8, t+4)
#pragma unroll
        for (int j0 = 0; j0 < NTL; j0 += JG) {
            uint32_t bh[JG][2], bl[JG][2];
#pragma unroll
            for (int j = 0; j < JG; ++j) {
                tf32_split(wl[(j0 + j) * 8 * LD + ks * 8], bh[j][0], bl[j][0]);
                tf32_split(wl[(j0 + j) * 8 * LD + ks * 8 + 4], bh[j][1], bl[j][1]);
            }
#if !defined(HP_TM_TERMS) || HP_TM_TERMS >= 3
#pragma unroll
            for (int j = 0; j < JG; ++j) mma_tf32(acc[j0 + j], al, bh[j][0], bh[j][1]);
#endif
#if !defined(HP_TM_TERMS) || HP_TM_TERMS >= 2
#pragma unroll
            for (int j = 0; j < JG; ++j) mma_tf32(acc[j0 + j], ah, bl[j][0], bl[j][1]);
#endif
#pragma unroll
            for (int j = 0; j < JG; ++j) mma_tf32(acc[j0 + j], ah, bh[j][0], bh[j][1]);
        }
    }
}

template <int NTL>
__device__ __forceinline__ void tm_init_bias(float (&acc)[NTL][4], const float *__restrict__ bias, int t) {
#pragma unroll
    for (int j = 0; j < NTL; ++j) {
        const float b0 = bias[8 * j + t], b1 = bias[8 * j + t + 4];
        acc[j][0] = b0, acc[j][1] = b1, acc[j][2] = b0, acc[j][3] = b1;
    }
}
template <int NTL>
__device__ __forceinline__ void tm_zero(float (&acc)[NTL][4]) {
#pragma unroll
    for (int j = 0; j < NTL; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
}
// ReLU in place; returns the gate bits (bit 4j + r)
template <int NTL>
__device__ __forceinline__ unsigned long long tm_relu(float (&v)[NTL][4]) {
    unsigned long long m = 0;
#pragma unroll
    for (int j = 0; j < NTL; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const bool on = v[j][r] > 0.f;
            v[j][r] = on ? v[j][r] : 0.f;
            m |= (unsigned long long)on << (4 * j + r);
        }
    return m;
}
template <int NTL>
__device__ __forceinline__ void tm_relu_only(float (&v)[NTL][4]) {
#pragma unroll
    for (int j = 0; j < NTL; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) v[j][r] = fmaxf(v[j][r], 0.f);
}
template <int NTL>
__device__ __forceinline__ void tm_gate(float (&v)[NTL][4], unsigned long long m) {
#pragma unroll
    for (int j = 0; j < NTL; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) v[j][r] = ((m >> (4 * j + r)) & 1ull) ? v[j][r] : 0.f;
}
// layer 1 (K = 3) on the FP32 pipe, in fragment layout.  w1p: [32][4] = (w0, w1, w2, bias)
__device__ __forceinline__ void tm_layer1(const float (&x0)[3], const float (&x1)[3], float (&a1)[4][4], const float *__restrict__ w1p, int t) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 wa = *reinterpret_cast<const float4 *>(w1p + (8 * j + t) * 4);
        const float4 wb = *reinterpret_cast<const float4 *>(w1p + (8 * j + t + 4) * 4);
        a1[j][0] = __fmaf_rn(wa.z, x0[2], __fmaf_rn(wa.y, x0[1], __fmaf_rn(wa.x, x0[0], wa.w)));
        a1[j][1] = __fmaf_rn(wb.z, x0[2], __fmaf_rn(wb.y, x0[1], __fmaf_rn(wb.x, x0[0], wb.w)));
        a1[j][2] = __fmaf_rn(wa.z, x1[2], __fmaf_rn(wa.y, x1[1], __fmaf_rn(wa.x, x1[0], wa.w)));
        a1[j][3] = __fmaf_rn(wb.z, x1[2], __fmaf_rn(wb.y, x1[1], __fmaf_rn(wb.x, x1[0], wb.w)));
    }
}

// ---- forward ------------------------------------------------------------------------------------------------------------------
// shared-memory map of ONE sample's weights (floats); two of these (current sample / next sample being prefetched)
constexpr int TMF_W1P = 0;                               // [32][4]
constexpr int TMF_W2 = TMF_W1P + C1 * 4;                 // [64][36]
constexpr int TMF_W3 = TMF_W2 + C2 * (C1 + 4);           // [128][68]
constexpr int TMF_W4 = TMF_W3 + C3 * (C2 + 4);           // [64][132]
constexpr int TMF_W5 = TMF_W4 + C4 * (C3 + 4);           // [8][68], rows 3..7 zero
constexpr int TMF_B = TMF_W5 + 8 * (C4 + 4);             // b2[64] b3[128] b4[64] b5[8]
constexpr int TMF_FLOATS = TMF_B + C2 + C3 + C4 + 8;
constexpr size_t TMF_SMEM = (size_t)2 * TMF_FLOATS * sizeof(float);
static_assert(TMF_SMEM <= 227 * 1024, "forward weights (double buffered) do not fit in shared memory");

// the parts of a weight buffer that do not depend on the sample: zero rows of W5, zero biases when the network has none
template <int NT>
__device__ __forceinline__ void tmf_fill_const(const TNArgs &a, float *__restrict__ S, int tid) {
    for (int i = tid; i < 5 * (C4 + 4); i += NT) S[TMF_W5 + 3 * (C4 + 4) + i] = 0.f;
    for (int i = tid; i < C1; i += NT) S[TMF_W1P + 4 * i + 3] = 0.f;
    for (int i = tid; i < C2 + C3 + C4 + 8; i += NT) S[TMF_B + i] = 0.f;
}
// stage sample b's weights into S with cp.async
template <int NT>
__device__ __forceinline__ void tmf_stage(const TNArgs &a, int b, float *__restrict__ S, int tid) {
    const float *wg = a.weights + (size_t)b * a.W;
    for (int i = tid; i < C1 * 4; i += NT) {
        const int o = i >> 2, c = i & 3;
        if (c < 3) cp_async4(S + TMF_W1P + i, wg + a.offw[0] + o * 3 + c);
        else if (a.offb[0] >= 0) cp_async4(S + TMF_W1P + i, wg + a.offb[0] + o);
    }
    tm_stage_nat<C1, C2, NT>(wg + a.offw[1], S + TMF_W2, tid);
    tm_stage_nat<C2, C3, NT>(wg + a.offw[2], S + TMF_W3, tid);
    tm_stage_nat<C3, C4, NT>(wg + a.offw[3], S + TMF_W4, tid);
    tm_stage_nat<C4, 3, NT>(wg + a.offw[4], S + TMF_W5, tid);
    for (int i = tid; i < C2 + C3 + C4 + 3; i += NT) {
        int l, o;
        if (i < C2) l = 1, o = i;
        else if (i < C2 + C3) l = 2, o = i - C2;
        else if (i < C2 + C3 + C4) l = 3, o = i - C2 - C3;
        else l = 4, o = i - C2 - C3 - C4;
        if (a.offb[l] >= 0) cp_async4(S + TMF_B + i, wg + a.offb[l] + o);
    }
}
// the mbarrier receives one arrival from this thread once all its cp.async so far have landed
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(TMF_THREADS, 1) tn_mma_forward_kernel(const TNArgs a) {
    extern __shared__ __align__(16) float sm[];
    __shared__ uint64_t bars[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3, sg = (g >> 1) + 4 * (g & 1);
    // flat list of (sample, 16-point unit), sample-major, split evenly over the CTAs; a CTA walks its range two samples at a time
    // (one weight buffer each).  Warps take units of the pair with a fixed stride and wait for the buffer a unit needs on that
    // buffer's mbarrier, so the second sample's weights stream in under the first sample's math and no warp waits for another.
    const int upn = (a.N + 15) >> 4;
    const long long TU = (long long)a.B * upn;
    const long long u0 = (long long)blockIdx.x * TU / gridDim.x, u1 = (long long)(blockIdx.x + 1) * TU / gridDim.x;
    if (u0 >= u1) return;
    const int bfirst = (int)(u0 / upn), blast = (int)((u1 - 1) / upn);
    if (tid == 0) mbar_init(&bars[0], TMF_THREADS), mbar_init(&bars[1], TMF_THREADS);
    tmf_fill_const<TMF_THREADS>(a, sm, tid);
    tmf_fill_const<TMF_THREADS>(a, sm + TMF_FLOATS, tid);
    __syncthreads();
    for (int bb = bfirst, round = 0; bb <= blast; bb += 2, ++round) {
        if (round > 0) __syncthreads();  // every warp is done with both buffers
        tmf_stage<TMF_THREADS>(a, bb, sm, tid);
        cp_async_mbar_arrive(&bars[0]);
        if (bb + 1 <= blast) {
            tmf_stage<TMF_THREADS>(a, bb + 1, sm + TMF_FLOATS, tid);
            cp_async_mbar_arrive(&bars[1]);
        }
        const long long s0 = u0 > (long long)bb * upn ? u0 : (long long)bb * upn;
        const long long s1 = u1 < (long long)(bb + 2) * upn ? u1 : (long long)(bb + 2) * upn;
        for (long long u = s0 + warp; u < s1; u += TMF_WARPS) {
            const int b = (int)(u / upn), which = b - bb;
            const float *S = sm + which * TMF_FLOATS;
            mbar_wait(&bars[which], round & 1);
            const float *pts = a.points + (size_t)b * a.pstride;
            const int r0 = (int)(u - (long long)b * upn) * 16 + g, r1 = r0 + 8;
            float x0[3] = {0.f, 0.f, 0.f}, x1[3] = {0.f, 0.f, 0.f};
            if (r0 < a.N) x0[0] = __ldg(pts + (size_t)r0 * 3), x0[1] = __ldg(pts + (size_t)r0 * 3 + 1), x0[2] = __ldg(pts + (size_t)r0 * 3 + 2);
            if (r1 < a.N) x1[0] = __ldg(pts + (size_t)r1 * 3), x1[1] = __ldg(pts + (size_t)r1 * 3 + 1), x1[2] = __ldg(pts + (size_t)r1 * 3 + 2);
            float a1[4][4];
            tm_layer1(x0, x1, a1, S + TMF_W1P, t);
            tm_relu_only<4>(a1);
            float a2[8][4];
            tm_init_bias<8>(a2, S + TMF_B, t);
            tm_layer<C1, C2>(a1, a2, S + TMF_W2 + sg * (C1 + 4) + t);
            tm_relu_only<8>(a2);
            float a3[16][4];
            tm_init_bias<16>(a3, S + TMF_B + C2, t);
            tm_layer<C2, C3>(a2, a3, S + TMF_W3 + sg * (C2 + 4) + t);
            tm_relu_only<16>(a3);
            float a4[8][4];
            tm_init_bias<8>(a4, S + TMF_B + C2 + C3, t);
            tm_layer<C3, C4>(a3, a4, S + TMF_W4 + sg * (C3 + 4) + t);
            tm_relu_only<8>(a4);
            float y[1][4];
            tm_init_bias<1>(y, S + TMF_B + C2 + C3 + C4, t);
            tm_layer<C4, 8>(a4, y, S + TMF_W5 + sg * (C4 + 4) + t);
            if (t < 3) {  // lane (g,t) holds coordinate t of rows r0 (y[0][0]) and r1 (y[0][2])
                if (a.channels_first) {
                    float *dst = a.out + ((size_t)b * 3 + t) * a.N;
                    if (r0 < a.N) dst[r0] = y[0][0];
                    if (r1 < a.N) dst[r1] = y[0][2];
                } else {
                    float *dst = a.out + (size_t)b * a.N * 3 + t;
                    if (r0 < a.N) dst[(size_t)r0 * 3] = y[0][0];
                    if (r1 < a.N) dst[(size_t)r1 * 3] = y[0][2];
                }
            }
        }
    }
}

// ---- backward -----------------------------------------------------------------------------------------------------------------
// One CTA of 12 warps per SM, 128-point tiles, two roles:
//   * warps 0-7, the CHAIN: forward recompute and dgrad of 16 points each, in registers (as the forward kernel).  They store the
//     tile's activations -- and later, in the same places, its pre-activation gradients Z_L -- to shared memory as [point][channel]
//     rows (stride 296 == 8 mod 32), only because wgrad contracts over POINTS and therefore needs them transposed;
//   * warps 8-11, the WGRAD helpers: dW_L[o][k] = sum_p Z_L[p][o] A_{L-1}[p][k] reads both operands straight from those rows as
//     conflict-free tensor-core fragments.  A tile's sum is accumulated on the tensor cores, then added in fp32 to the sample's
//     running gradient, which lives in TENSOR MEMORY (144 columns of the helper's own lanes, tcgen05.ld / tcgen05.st): no persistent
//     registers, so all 12 warps fit at 168 registers.  The helpers also own the small gradients (biases, the 3-wide layers).
// Per tile every warp issues the same number of MMAs (1728: chain 864 + 864, helper 768 + 768 + 192), three warps per scheduler.
// Hand-over of a block of rows between the roles is a pair of named barriers (the writer arrives on FULL after its stores, the reader
// syncs on it; the reader arrives on EMPTY when it is done, the writer syncs on it before it overwrites the block in place).
constexpr int TMB_CHAIN_WARPS = 8, TMB_CHAIN_THREADS = 256;
constexpr int TMB_HELP_THREADS = 128;
constexpr int TMB_ALL_THREADS = TMB_CHAIN_THREADS + TMB_HELP_THREADS;  // 384
constexpr int TM_ACT_LD = 296;                // activation row stride: 32 + 64 + 128 + 64 + 8, == 8 (mod 32)
constexpr int TM_A1 = 0, TM_A2 = 32, TM_A3 = 96, TM_A4 = 224;  // column of each layer's block inside a row
constexpr int TMB_ACT = 0;                                  // [128][296]
constexpr int TMB_W2 = TMB_ACT + TN_T * TM_ACT_LD;          // [64][32]   the sample's matrices, dense and XOR-swizzled (tm_layer_sw),
constexpr int TMB_W3 = TMB_W2 + C2 * C1;                    // [128][64]  resident for all tiles of the sample: ONE copy serves the
constexpr int TMB_W4 = TMB_W3 + C3 * C2;                    // [64][128]  forward (along rows) and dgrad (down columns)
constexpr int TMB_XS = TMB_W4 + C4 * C3;                    // [128][4] (x, y, z, 0)
constexpr int TMB_GS = TMB_XS + TN_T * 4;                   // [128][4] upstream gradient of the 3 output coordinates
constexpr int TMB_W1P = TMB_GS + TN_T * 4;                  // [32][4]
constexpr int TMB_W5 = TMB_W1P + C1 * 4;                    // [3][64]
constexpr int TMB_B = TMB_W5 + 3 * C4;                      // b2[64] b3[128] b4[64]
constexpr int TMB_FLOATS = TMB_B + C2 + C3 + C4;
constexpr size_t TMB_SMEM = (size_t)TMB_FLOATS * sizeof(float);
static_assert(TMB_SMEM + 64 <= 227 * 1024, "backward tile does not fit in shared memory");
// named barriers (0 is __syncthreads)
enum : int { TMB_BAR_CHAIN = 1, TMB_BAR_TILE_EMPTY, TMB_BAR_A4_FULL, TMB_BAR_A4_EMPTY, TMB_BAR_Z4_FULL, TMB_BAR_A3_EMPTY,
             TMB_BAR_Z3_FULL, TMB_BAR_A2_EMPTY, TMB_BAR_Z2_FULL, TMB_BAR_A1_EMPTY, TMB_BAR_Z1_FULL, TMB_BAR_HELPERS };

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) {
    __threadfence_block();
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---- tensor memory as thread-private storage: each thread of a warp owns the columns of its own lane ----
#define HP_R16(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), \
                     "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]), "=r"(v[o + 12]),            \
                     "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])
#define HP_I16(v, o) "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]),        \
                     "r"(v[o + 7]), "r"(v[o + 8]), "r"(v[o + 9]), "r"(v[o + 10]), "r"(v[o + 11]), "r"(v[o + 12]), "r"(v[o + 13]),   \
                     "r"(v[o + 14]), "r"(v[o + 15])
// N (a multiple of 16) consecutive columns of this thread's lane <-> registers
template <int N>
__device__ __forceinline__ void tmem_load(uint32_t taddr, uint32_t (&v)[N]) {
#pragma unroll
    for (int o = 0; o < N; o += 16)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : HP_R16(v, o)
                     : "r"(taddr + o));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tmem_store(uint32_t taddr, const uint32_t (&v)[N]) {
#pragma unroll
    for (int o = 0; o < N; o += 16)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr + o),
                     HP_I16(v, o)
                     : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// columns [taddr, taddr + 4*TILES) += acc (the tile's tensor-core sum joins the sample's running fp32 gradient)
template <int TILES>
__device__ __forceinline__ void tmem_accumulate(uint32_t taddr, const float (&acc)[TILES][4]) {
#pragma unroll
    for (int h = 0; h < TILES * 4; h += 16) {
        uint32_t v[16];
        tmem_load<16>(taddr + h, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + acc[(h + i) >> 2][(h + i) & 3]);
        tmem_store<16>(taddr + h, v);
    }
}

// fragment -> shared rows; row0 = act + (16*warp + g)*LD + column of the layer + t
template <int NTL>
__device__ __forceinline__ void tm_store_rows(float *__restrict__ row0, const float (&v)[NTL][4]) {
#pragma unroll
    for (int j = 0; j < NTL; ++j) {
        row0[8 * j] = v[j][0], row0[8 * j + 4] = v[j][1];
        row0[8 * TM_ACT_LD + 8 * j] = v[j][2], row0[8 * TM_ACT_LD + 8 * j + 4] = v[j][3];
    }
}
// dW tile on the tensor cores: acc[mi*NT+ni] += sum over the tile's 128 points of Z[p][zc + 16mi + g (+8)] * A[p][ac + 8ni + 2t (+1)]
template <int MT, int NT, int PTS = TN_T>
__device__ __forceinline__ void tm_wgrad(const float *__restrict__ act, int zc, int ac, float (&acc)[MT * NT][4], int g, int t) {
#ifdef HP_TMB_NO_WGRAD
    return;
#endif
    const float *zr = act + t * TM_ACT_LD + zc + g;
    const float *ar = act + t * TM_ACT_LD + ac + g;
#pragma unroll(MT * NT >= 16 ? 1 : 4)
    for (int p0 = 0; p0 < PTS; p0 += 8) {
        uint32_t ah[MT][4], al[MT][4];
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
            tf32_split(zr[p0 * TM_ACT_LD + 16 * mi], ah[mi][0], al[mi][0]);                      // (o = g,   p = t)
            tf32_split(zr[p0 * TM_ACT_LD + 16 * mi + 8], ah[mi][1], al[mi][1]);                  // (o = g+8, p = t)
            tf32_split(zr[(p0 + 4) * TM_ACT_LD + 16 * mi], ah[mi][2], al[mi][2]);                // (o = g,   p = t+4)
            tf32_split(zr[(p0 + 4) * TM_ACT_LD + 16 * mi + 8], ah[mi][3], al[mi][3]);            // (o = g+8, p = t+4)
        }
        uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) {
            tf32_split(ar[p0 * TM_ACT_LD + 8 * ni], bh[ni][0], bl[ni][0]);        // (p = t,   k = g)
            tf32_split(ar[(p0 + 4) * TM_ACT_LD + 8 * ni], bh[ni][1], bl[ni][1]);  // (p = t+4, k = g)
        }
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) mma_tf32(acc[mi * NT + ni], al[mi], bh[ni][0], bh[ni][1]);
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) mma_tf32(acc[mi * NT + ni], ah[mi], bl[ni][0], bl[ni][1]);
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) mma_tf32(acc[mi * NT + ni], ah[mi], bh[ni][0], bh[ni][1]);
    }
}
// running gradient (tensor memory) -> dW[O][K] (flat): rows o0 + 16mi + g (+8), columns k0 + 8ni + 2t (+1); the columns are cleared
template <int MT, int NT, int K>
__device__ __forceinline__ void tm_flush_dw(uint32_t taddr, float *__restrict__ dst, int o0, int k0, int g, int t,
                                            const float *__restrict__ addend = nullptr) {
#pragma unroll
    for (int h = 0; h < MT * NT * 4; h += 16) {
        uint32_t v[16];
        tmem_load<16>(taddr + h, v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int tile = (h >> 2) + q, mi = tile / NT, ni = tile - mi * NT;
            const int off = (o0 + 16 * mi + g) * K + k0 + 8 * ni + 2 * t;
            float x0 = __uint_as_float(v[4 * q]), x1 = __uint_as_float(v[4 * q + 1]), x2 = __uint_as_float(v[4 * q + 2]),
                  x3 = __uint_as_float(v[4 * q + 3]);
            if (addend) x0 += addend[off], x1 += addend[off + 1], x2 += addend[off + 8 * K], x3 += addend[off + 8 * K + 1];
            dst[off] = x0, dst[off + 1] = x1;
            dst[off + 8 * K] = x2, dst[off + 8 * K + 1] = x3;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0u;
        tmem_store<16>(taddr + h, v);
    }
}
// column sum over the tile's rows
template <int PTS = TN_T>
__device__ __forceinline__ float tm_col_sum(const float *__restrict__ col) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
    for (int p = 0; p < PTS; p += 4) {
        s0 += col[p * TM_ACT_LD], s1 += col[(p + 1) * TM_ACT_LD];
        s2 += col[(p + 2) * TM_ACT_LD], s3 += col[(p + 3) * TM_ACT_LD];
    }
    return (s0 + s1) + (s2 + s3);
}

// ---- one swizzled copy of W[O][K] for both contractions: element (r, c) lives at r*K + (c ^ 4*(r & 7)) ----
template <int K, int O, int NT>
__device__ __forceinline__ void tm_stage_sw(const float *__restrict__ Wg, float *__restrict__ dst, int tid) {
    for (int i = tid; i < O * K; i += NT) {
        const int o = i / K, k = i - o * K;
        cp_async4(dst + o * K + (k ^ (4 * (o & 7))), Wg + i);
    }
}
// tm_layer on the swizzled copy.  KC = contraction length, N = produced channels.
//   DGRAD = false: W is [N][KC], the contraction runs along a row: fragment rows 8j + s(g), columns 8ks + t (+4): conflict-free;
//   DGRAD = true:  W is [KC][N], the contraction runs down a column: rows 8ks + t (+4), columns 8j + s(g): two lanes per bank.
template <int KC, int N, bool DGRAD>
__device__ __forceinline__ void tm_layer_sw(const float (&in)[KC / 8][4], float (&acc)[N / 8][4], const float *__restrict__ W, int sg, int t) {
    constexpr int NTL = N / 8, JG = NTL >= HP_TM_JG ? HP_TM_JG : NTL;
    static_assert(NTL >= 4 && NTL % 4 == 0, "channel tiles");
    // FWD: row s(g) of a tile, swizzle 4*s(g): bit 2 moves t into t or t+4, bits 3-4 flip the contraction step (per ks, below)
    const float *f0 = W + sg * KC + t + 4 * (sg & 1), *f1 = W + sg * KC + t + 4 * (1 - (sg & 1));
    const int fx = 8 * (sg >> 1);
    // DGRAD: row 8ks + t (+4), swizzle 4*t (+16): bit 2 flips s(g), bit 3 swaps even/odd channel tiles, bit 4 (second row) tiles j, j^2
    const float *dbase = W + t * N + (sg ^ (4 * (t & 1)));
    const float *dE = dbase + 8 * (t >> 1), *dO = dbase - 8 * (t >> 1);
#pragma unroll
    for (int ks = 0; ks < KC / 8; ++ks) {
        uint32_t ah[4], al[4];
        tf32_split(in[ks][0], ah[0], al[0]);  // a0 = (g,   t)
        tf32_split(in[ks][2], ah[1], al[1]);  // a1 = (g+8, t)
        tf32_split(in[ks][1], ah[2], al[2]);  // a2 = (g,   t+4)
        tf32_split(in[ks][3], ah[3], al[3]);  // a3 = (g+8, t+4)
        const int fo = (8 * ks) ^ fx;
#pragma unroll
        for (int j0 = 0; j0 < NTL; j0 += JG) {
            uint32_t bh[JG][2], bl[JG][2];
#pragma unroll
            for (int j = 0; j < JG; ++j) {
                const int jj = j0 + j;
                float w0, w1;
                if (!DGRAD) {
                    w0 = f0[jj * 8 * KC + fo], w1 = f1[jj * 8 * KC + fo];
                } else {
                    const float *d = (jj & 1) ? dO : dE;
                    w0 = d[(8 * ks) * N + 8 * jj], w1 = d[(8 * ks + 4) * N + 8 * (jj ^ 2)];
                }
                tf32_split(w0, bh[j][0], bl[j][0]);
                tf32_split(w1, bh[j][1], bl[j][1]);
            }
#pragma unroll
            for (int j = 0; j < JG; ++j) mma_tf32(acc[j0 + j], al, bh[j][0], bh[j][1]);
#pragma unroll
            for (int j = 0; j < JG; ++j) mma_tf32(acc[j0 + j], ah, bl[j][0], bl[j][1]);
#pragma unroll
            for (int j = 0; j < JG; ++j) mma_tf32(acc[j0 + j], ah, bh[j][0], bh[j][1]);
        }
    }
}

#ifdef HP_TM_TRACE
__device__ unsigned long long g_tm_cta[256][4];  // per CTA: kernel entry, first tile, after the last tile, exit
__device__ __forceinline__ void tm_cta_stamp(int ev) {
    if (threadIdx.x == 0) {
        unsigned long long tns;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
        g_tm_cta[blockIdx.x][ev] = tns;
    }
}
#define TM_CTA(ev) tm_cta_stamp(ev)
__device__ unsigned long long g_tm_trace[2][64][16];  // [role][tile][event] %globaltimer stamps of CTA 0 (chain warp 0 / helper warp 8)
__device__ __forceinline__ void tm_trace(int role, int tile, int ev, int lane) {
    if (blockIdx.x == 0 && lane == 0 && tile < 64) {
        unsigned long long tns;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
        g_tm_trace[role][tile][ev] = tns;
    }
}
#define TM_TRACE(role, ev) do { if ((role) == 0 ? warp == 0 : warp == TMB_CHAIN_WARPS) tm_trace(role, (int)(f - f0), ev, lane); } while (0)
#else
#define TM_TRACE(role, ev) do { } while (0)
#define TM_CTA(ev) do { } while (0)
#endif

template <bool GRAD_POINTS>
__global__ void __launch_bounds__(TMB_ALL_THREADS, 1) tn_mma_backward_kernel(const TNArgs a) {
    extern __shared__ __align__(16) float sm[];
    float *act = sm + TMB_ACT, *W2s = sm + TMB_W2, *W3s = sm + TMB_W3, *W4s = sm + TMB_W4, *xs = sm + TMB_XS, *gs = sm + TMB_GS;
    float *w1p = sm + TMB_W1P, *w5 = sm + TMB_W5, *bs = sm + TMB_B;
    __shared__ int is_last;
    __shared__ uint32_t tmem_base_slot;
    unsigned int *slot_of = reinterpret_cast<unsigned int *>(act);  // [TN_MAX_GRID], only inside flush (no tile is live then)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3, sg = (g >> 1) + 4 * (g & 1);
    const bool helper = warp >= TMB_CHAIN_WARPS;
    const int ht = tid - TMB_CHAIN_THREADS, hw = warp - TMB_CHAIN_WARPS;  // helper thread / warp index
    const int ntiles = (a.N + TN_T - 1) / TN_T;
    const long long TT = (long long)a.B * ntiles;
    const long long G = gridDim.x;
    const long long f0 = (long long)blockIdx.x * TT / G, f1 = (long long)(blockIdx.x + 1) * TT / G;
    const int first_sample = (int)(f0 / ntiles);
    TM_CTA(0);

    // tensor memory: 256 columns; helper warp hw owns lanes 32*hw.. of them (columns 0-63 dW4, 64-127 dW3, 128-143 dW2)
    if (warp == TMB_CHAIN_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base_slot + ((uint32_t)(32 * (hw & 3)) << 16);
    // helpers: small gradients, a few elements per thread
    //   sC0..2: dW5[0..2][ht] (ht < 64) | sB: db3[ht] | sA0: db4[ht] (ht < 64), db2[ht - 64] | sA1: db1[ht - 96] (ht >= 96), db5[ht - 64] (64 <= ht < 67)
    //   sD: dW1[ht] (ht < 96)
    float sC0 = 0.f, sC1 = 0.f, sC2 = 0.f, sB = 0.f, sA0 = 0.f, sA1 = 0.f, sD = 0.f;
    if (helper) {
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
        for (int c = 0; c < 144; c += 16) tmem_store<16>(tm + c, z);
    }
    // Write the sample's gradient (helpers) and, when other CTAs also hold tiles of it, fold the partials (everybody).
    auto flush = [&](int b) {
        const long long s0 = (long long)b * ntiles, s1 = s0 + ntiles - 1;
        const int c_lo = (int)(((s0 + 1) * G - 1) / TT), c_hi = (int)(((s1 + 1) * G - 1) / TT);
        const bool shared_sample = c_lo != c_hi;
        if (helper) {
            float *dst = shared_sample ? a.partial + ((size_t)blockIdx.x * a.S + (b - first_sample)) * a.W : a.gweights + (size_t)b * a.W;
            tm_flush_dw<4, 4, C3>(tm, dst + a.offw[3], 0, hw * 32, g, t);
            tm_flush_dw<4, 4, C2>(tm + 64, dst + a.offw[2], (hw & 1) * 64, (hw >> 1) * 32, g, t);
            tm_flush_dw<4, 1, C1>(tm + 128, dst + a.offw[1], 0, hw * 8, g, t);
            if (ht < 64) dst[a.offw[4] + ht] = sC0, dst[a.offw[4] + 64 + ht] = sC1, dst[a.offw[4] + 128 + ht] = sC2;
            if (a.offb[2] >= 0) dst[a.offb[2] + ht] = sB;
            if (ht < 64) { if (a.offb[3] >= 0) dst[a.offb[3] + ht] = sA0; }
            else { if (a.offb[1] >= 0) dst[a.offb[1] + ht - 64] = sA0; }
            if (ht >= 96) { if (a.offb[0] >= 0) dst[a.offb[0] + ht - 96] = sA1; }
            else if (ht >= 64 && ht < 67) { if (a.offb[4] >= 0) dst[a.offb[4] + ht - 64] = sA1; }
            if (ht < 96) dst[a.offw[0] + ht] = sD;
            sC0 = sC1 = sC2 = sB = sA0 = sA1 = sD = 0.f;
        }
        if (shared_sample) {
            __threadfence();
            __syncthreads();
            if (tid == 0) is_last = (atomicAdd(a.counters + b, 1u) == (unsigned)(c_hi - c_lo));
            __syncthreads();
            if (is_last) {
                __threadfence();
                const int ncontrib = c_hi - c_lo + 1;
                for (int q = tid; q < ncontrib; q += TMB_ALL_THREADS) {
                    const int c = c_lo + q;
                    const int fs = (int)(((long long)c * TT / G) / ntiles);
                    slot_of[q] = (unsigned)(c * a.S + (b - fs));
                }
                __syncthreads();
                float *gw = a.gweights + (size_t)b * a.W;
                tn_fold_partials<TMB_ALL_THREADS>(a.partial, slot_of, ncontrib, a.W, gw, tid);
            }
            __syncthreads();
        }
    };

    int cur = -1;
    TM_CTA(1);
    if (!helper) {
        // =========================================== CHAIN ===========================================
        for (long long f = f0; f < f1; ++f) {
            const int b = (int)(f / ntiles), tl = (int)(f - (long long)b * ntiles);
            const float *wg = a.weights + (size_t)b * a.W;
            TM_TRACE(0, 0);
            const int n0 = tl * TN_T;
            // this warp's 16 points and their upstream gradients: loaded before the wait below, stored (the helpers read them too) after it
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), gv = xv;
            if (lane < 16) {
                const int p = n0 + 16 * warp + lane;
                if (p < a.N) {
                    const float *px = a.points + (size_t)b * a.pstride + (size_t)p * 3;
                    xv.x = __ldg(px), xv.y = __ldg(px + 1), xv.z = __ldg(px + 2);
                    if (a.channels_first) {
                        const float *pg = a.gout + (size_t)b * 3 * a.N + p;
                        gv.x = __ldg(pg), gv.y = __ldg(pg + a.N), gv.z = __ldg(pg + 2 * (size_t)a.N);
                    } else {
                        const float *pg = a.gout + ((size_t)b * a.N + p) * 3;
                        gv.x = __ldg(pg), gv.y = __ldg(pg + 1), gv.z = __ldg(pg + 2);
                    }
                }
            }
            bool waited = (f == f0);
            if (b != cur) {  // new sample: everything of the previous tile must be over before the weights change
                if (!waited) named_sync(TMB_BAR_TILE_EMPTY, TMB_ALL_THREADS), waited = true;
                if (cur >= 0) flush(cur);
                // every chain warp is past the previous tile (TILE_EMPTY / flush): restage the sample's weights
                tm_stage_sw<C1, C2, TMB_CHAIN_THREADS>(wg + a.offw[1], W2s, tid);
                tm_stage_sw<C2, C3, TMB_CHAIN_THREADS>(wg + a.offw[2], W3s, tid);
                tm_stage_sw<C3, C4, TMB_CHAIN_THREADS>(wg + a.offw[3], W4s, tid);
                cp_async_commit();
                for (int i = tid; i < C1 * 4; i += TMB_CHAIN_THREADS) {
                    const int o = i >> 2, c = i & 3;
                    w1p[i] = c < 3 ? __ldg(wg + a.offw[0] + o * 3 + c) : (a.offb[0] >= 0 ? __ldg(wg + a.offb[0] + o) : 0.f);
                }
                for (int i = tid; i < 3 * C4; i += TMB_CHAIN_THREADS) w5[i] = __ldg(wg + a.offw[4] + i);
                for (int i = tid; i < C2 + C3 + C4; i += TMB_CHAIN_THREADS) {
                    int l, o;
                    if (i < C2) l = 1, o = i;
                    else if (i < C2 + C3) l = 2, o = i - C2;
                    else l = 3, o = i - C2 - C3;
                    bs[i] = a.offb[l] >= 0 ? __ldg(wg + a.offb[l] + o) : 0.f;
                }
                cp_async_wait<0>();
                named_sync(TMB_BAR_CHAIN, TMB_CHAIN_THREADS);
                cur = b;
            }
            TM_TRACE(0, 1);
            // ---- layers 1 and 2 in registers while the helpers finish the previous tile (its rows are still theirs) ----
            unsigned int m1, m2, m4;
            unsigned long long m3;
            float a2[8][4];
            float a1[4][4];
            {
                const float x0[3] = {__shfl_sync(0xffffffffu, xv.x, g), __shfl_sync(0xffffffffu, xv.y, g), __shfl_sync(0xffffffffu, xv.z, g)};
                const float x1[3] = {__shfl_sync(0xffffffffu, xv.x, g + 8), __shfl_sync(0xffffffffu, xv.y, g + 8),
                                     __shfl_sync(0xffffffffu, xv.z, g + 8)};
                tm_layer1(x0, x1, a1, w1p, t);
            }
            m1 = (unsigned int)tm_relu<4>(a1);
            tm_init_bias<8>(a2, bs, t);
            tm_layer_sw<C1, C2, false>(a1, a2, W2s, sg, t);
            m2 = (unsigned int)tm_relu<8>(a2);
            if (!waited) named_sync(TMB_BAR_TILE_EMPTY, TMB_ALL_THREADS);  // the helpers are done with the previous tile
            TM_TRACE(0, 2);
            float *row0 = act + (16 * warp + g) * TM_ACT_LD + t;
            if (lane < 16) {
                *reinterpret_cast<float4 *>(xs + (16 * warp + lane) * 4) = xv;
                *reinterpret_cast<float4 *>(gs + (16 * warp + lane) * 4) = gv;
            }
            __syncwarp();
            tm_store_rows<4>(row0 + TM_A1, a1);
            tm_store_rows<8>(row0 + TM_A2, a2);
            float z4[8][4];
            {
                float a4[8][4];
                {
                    float a3[16][4];
                    tm_init_bias<16>(a3, bs + C2, t);
                    tm_layer_sw<C2, C3, false>(a2, a3, W3s, sg, t);
                    m3 = tm_relu<16>(a3);
                    tm_store_rows<16>(row0 + TM_A3, a3);
                    tm_init_bias<8>(a4, bs + C2 + C3, t);
                    tm_layer_sw<C3, C4, false>(a3, a4, W4s, sg, t);
                }
                m4 = (unsigned int)tm_relu<8>(a4);
                tm_store_rows<8>(row0 + TM_A4, a4);
            }
            TM_TRACE(0, 3);
            named_arrive(TMB_BAR_A4_FULL, TMB_ALL_THREADS);  // A1..A4, xs, gs of this warp's rows are in place
            // ---- layer 5: Z5 = dY.  Z4 = gate4 * (dY W5) in fragment layout, K = 3 on the FP32 pipe ----
            {
                const float4 g0 = *reinterpret_cast<const float4 *>(gs + (16 * warp + g) * 4);
                const float4 g1 = *reinterpret_cast<const float4 *>(gs + (16 * warp + g + 8) * 4);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int ca = 8 * j + t, cb = ca + 4;
                    const float wa0 = w5[ca], wa1 = w5[C4 + ca], wa2 = w5[2 * C4 + ca];
                    const float wc0 = w5[cb], wc1 = w5[C4 + cb], wc2 = w5[2 * C4 + cb];
                    z4[j][0] = __fmaf_rn(g0.z, wa2, __fmaf_rn(g0.y, wa1, g0.x * wa0));
                    z4[j][1] = __fmaf_rn(g0.z, wc2, __fmaf_rn(g0.y, wc1, g0.x * wc0));
                    z4[j][2] = __fmaf_rn(g1.z, wa2, __fmaf_rn(g1.y, wa1, g1.x * wa0));
                    z4[j][3] = __fmaf_rn(g1.z, wc2, __fmaf_rn(g1.y, wc1, g1.x * wc0));
                }
                tm_gate<8>(z4, m4);
            }
            TM_TRACE(0, 4);
            named_sync(TMB_BAR_A4_EMPTY, TMB_ALL_THREADS);  // the helpers have read A4 (dW5)
            TM_TRACE(0, 5);
            tm_store_rows<8>(row0 + TM_A4, z4);
            named_arrive(TMB_BAR_Z4_FULL, TMB_ALL_THREADS);
            // ---- layer 4: Z3 = gate3 * (Z4 W4) while the helpers take dW4 = Z4^T A3 ----
            float z2[8][4];
            {
                float z3[16][4];
                tm_zero<16>(z3);
                tm_layer_sw<C4, C3, true>(z4, z3, W4s, sg, t);
                tm_gate<16>(z3, m3);
                TM_TRACE(0, 6);
                named_sync(TMB_BAR_A3_EMPTY, TMB_ALL_THREADS);  // the helpers have read A3 (and Z4)
                TM_TRACE(0, 7);
                tm_store_rows<16>(row0 + TM_A3, z3);
                named_arrive(TMB_BAR_Z3_FULL, TMB_ALL_THREADS);
                // ---- layer 3: Z2 = gate2 * (Z3 W3) ----
                tm_zero<8>(z2);
                tm_layer_sw<C3, C2, true>(z3, z2, W3s, sg, t);
            }
            tm_gate<8>(z2, m2);
            TM_TRACE(0, 8);
            named_sync(TMB_BAR_A2_EMPTY, TMB_ALL_THREADS);
            TM_TRACE(0, 9);
            tm_store_rows<8>(row0 + TM_A2, z2);
            named_arrive(TMB_BAR_Z2_FULL, TMB_ALL_THREADS);
            // ---- layer 2: Z1 = gate1 * (Z2 W2) ----
            {
                float z1[4][4];
                tm_zero<4>(z1);
                tm_layer_sw<C2, C1, true>(z2, z1, W2s, sg, t);
                tm_gate<4>(z1, m1);
                TM_TRACE(0, 10);
                named_sync(TMB_BAR_A1_EMPTY, TMB_ALL_THREADS);
                TM_TRACE(0, 11);
                tm_store_rows<4>(row0 + TM_A1, z1);
            }
            named_arrive(TMB_BAR_Z1_FULL, TMB_ALL_THREADS);
            TM_TRACE(0, 12);
        }
    } else {
        // =========================================== WGRAD HELPERS ===========================================
        for (long long f = f0; f < f1; ++f) {
            const int b = (int)(f / ntiles), tl = (int)(f - (long long)b * ntiles);
            if (b != cur) {
                if (cur >= 0) flush(cur);
                cur = b;
            }
            const int n0 = tl * TN_T;
            // ---- layer 5: dW5[c][k] += sum_p dY[p][c] A4[p][k], db5 ----
            TM_TRACE(1, 0);
            named_sync(TMB_BAR_A4_FULL, TMB_ALL_THREADS);
            TM_TRACE(1, 1);
            if (ht < 64) {  // dW5[0..2][ht]
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, u0 = 0.f, u1 = 0.f, u2 = 0.f;
#pragma unroll 4
                for (int p = 0; p < TN_T; p += 2) {
                    const float av = act[p * TM_ACT_LD + TM_A4 + ht], aw = act[(p + 1) * TM_ACT_LD + TM_A4 + ht];
                    const float4 ga = *reinterpret_cast<const float4 *>(gs + p * 4), gb = *reinterpret_cast<const float4 *>(gs + p * 4 + 4);
                    s0 = __fmaf_rn(ga.x, av, s0), s1 = __fmaf_rn(ga.y, av, s1), s2 = __fmaf_rn(ga.z, av, s2);
                    u0 = __fmaf_rn(gb.x, aw, u0), u1 = __fmaf_rn(gb.y, aw, u1), u2 = __fmaf_rn(gb.z, aw, u2);
                }
                sC0 += s0 + u0, sC1 += s1 + u1, sC2 += s2 + u2;
            } else if (ht < 67) {  // db5
                float s = 0.f, u = 0.f;
#pragma unroll 4
                for (int p = 0; p < TN_T; p += 2) s += gs[p * 4 + ht - 64], u += gs[p * 4 + 4 + ht - 64];
                sA1 += s + u;
            }
            named_arrive(TMB_BAR_A4_EMPTY, TMB_ALL_THREADS);
            // ---- layer 4: dW4 += Z4^T A3, db4 ----
            TM_TRACE(1, 2);
            named_sync(TMB_BAR_Z4_FULL, TMB_ALL_THREADS);
            TM_TRACE(1, 3);
            {
                float acc[16][4];
                tm_zero<16>(acc);
                tm_wgrad<4, 4>(act, TM_A4, TM_A3 + hw * 32, acc, g, t);
                if (ht < 64) sA0 += tm_col_sum(act + TM_A4 + ht);
                named_arrive(TMB_BAR_A3_EMPTY, TMB_ALL_THREADS);
                tmem_accumulate<16>(tm, acc);
            }
            // ---- layer 3: dW3 += Z3^T A2, db3 ----
            TM_TRACE(1, 4);
            named_sync(TMB_BAR_Z3_FULL, TMB_ALL_THREADS);
            TM_TRACE(1, 5);
            {
                float acc[16][4];
                tm_zero<16>(acc);
                tm_wgrad<4, 4>(act, TM_A3 + (hw & 1) * 64, TM_A2 + (hw >> 1) * 32, acc, g, t);
                sB += tm_col_sum(act + TM_A3 + ht);
                named_arrive(TMB_BAR_A2_EMPTY, TMB_ALL_THREADS);
                tmem_accumulate<16>(tm + 64, acc);
            }
            // ---- layer 2: dW2 += Z2^T A1, db2 ----
            TM_TRACE(1, 6);
            named_sync(TMB_BAR_Z2_FULL, TMB_ALL_THREADS);
            TM_TRACE(1, 7);
            {
                float acc[4][4];
                tm_zero<4>(acc);
                tm_wgrad<4, 1>(act, TM_A2, TM_A1 + hw * 8, acc, g, t);
                if (ht >= 64) sA0 += tm_col_sum(act + TM_A2 + ht - 64);
                named_arrive(TMB_BAR_A1_EMPTY, TMB_ALL_THREADS);
                tmem_accumulate<4>(tm + 128, acc);
            }
            // ---- layer 1: dW1 += Z1^T X, db1, optionally dX = Z1 W1 ----
            TM_TRACE(1, 8);
            named_sync(TMB_BAR_Z1_FULL, TMB_ALL_THREADS);
            TM_TRACE(1, 9);
            if (ht < 96) {
                const int o = ht / 3, c = ht - o * 3;
                float s = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll 2
                for (int p = 0; p < TN_T; p += 4) {
                    s = __fmaf_rn(act[p * TM_ACT_LD + TM_A1 + o], xs[p * 4 + c], s);
                    s2 = __fmaf_rn(act[(p + 1) * TM_ACT_LD + TM_A1 + o], xs[p * 4 + 4 + c], s2);
                    s3 = __fmaf_rn(act[(p + 2) * TM_ACT_LD + TM_A1 + o], xs[p * 4 + 8 + c], s3);
                    s4 = __fmaf_rn(act[(p + 3) * TM_ACT_LD + TM_A1 + o], xs[p * 4 + 12 + c], s4);
                }
                sD += (s + s2) + (s3 + s4);
            }
            if (ht >= 96) sA1 += tm_col_sum(act + TM_A1 + ht - 96);  // db1 on the fourth helper warp, beside dW1 on the other three
            if (GRAD_POINTS) {
                for (int i = ht; i < 3 * TN_T; i += TMB_HELP_THREADS) {
                    const int p = i / 3, c = i - p * 3;
                    float sx = 0.f;
#pragma unroll 8
                    for (int o = 0; o < C1; ++o) sx = __fmaf_rn(act[p * TM_ACT_LD + TM_A1 + o], w1p[o * 4 + c], sx);
                    if (n0 + p < a.N) a.gpoints[((size_t)b * a.N + n0 + p) * 3 + c] = sx;
                }
            }
            TM_TRACE(1, 10);
            if (f + 1 < f1) named_arrive(TMB_BAR_TILE_EMPTY, TMB_ALL_THREADS);
        }
    }
    TM_CTA(2);
    if (cur >= 0) flush(cur);
    TM_CTA(3);
    // tensor memory goes back once every helper has read its columns
    if (helper) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        named_sync(TMB_BAR_HELPERS, TMB_HELP_THREADS);
        if (warp == TMB_CHAIN_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base_slot) : "memory");
    }
}

}  // namespace hp
