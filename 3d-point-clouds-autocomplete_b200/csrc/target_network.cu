// target_network.cu -- fused per-sample TargetNetwork MLP (forward and backward) for sm_100a.
//
// Replaces model/target_network.py:5-45 driven by the per-sample Python loop of
// model/full_model.py:67-74 (per sample: 5x torch.mm + bias + ReLU, ~15 launches) with ONE launch per
// batch.  Every sample b has its own flat weight vector weights[b] emitted by the hypernetwork, laid out
// per layer as W[out][in] row-major followed by b[out] (target_network.py:40-45, hyper_network.py:32-36).
//
// Fast path (all sample configs of the reference, settings/*.json.sample "layer_out_channels"):
//   3 -> 32 -> 64 -> 128 -> 64 -> 3, ReLU between layers, none on the output.
//   * a CTA (512 threads) owns one sample and walks tiles of 128 points of it;
//   * activations of a tile live in shared memory as [channel][point] rows (row stride 132 floats: every
//     LDS.128/STS.128 below is bank-conflict free) and in registers as 4-point x 8/4/2-channel micro tiles;
//   * weights are staged from the flat vector into shared memory as [out][in+4] rows; FORWARD keeps all five
//     layers resident for the CTA's life, BACKWARD streams them through a double buffer, each staging
//     overlapped with the previous layer's math;
//   * K = 3..128 contractions in FP32 FFMA (1e-5 parity with torch.mm fp32 rules out TF32/BF16 tensor cores);
//     a warp's 32 lanes form an 8 (point groups) x 4 (channel groups) grid, so each LDS.128 is one
//     shared-memory wavefront and feeds 16 FFMAs per lane;
//   * only the 3 output coordinates per point go to HBM (optionally straight into the trainer's [B,3,N]
//     layout, full_model.py:68,74).  No intermediate activation is ever written to HBM.
//   BACKWARD recomputes the forward of its tile (nothing was stashed), then walks the layers in reverse:
//     wgrad  dW_L[o][k] += sum_p Z_L[o][p] A_{L-1}[k][p]   accumulators persistent in REGISTERS across all
//                                                           tiles of the CTA (36 per thread),
//     dgrad  Z_{L-1}[k][p] = (A_{L-1}[k][p] > 0) * sum_o Z_L[o][p] W_L[o][k]   written in place of A_{L-1}.
//   Each CTA writes its partial dW once; the last CTA of a sample folds the S partials in ascending order
//   => bitwise deterministic, no float atomics.
// Generic path (any other widths / depth): simple shared-memory kernels, same semantics, not tuned.
#include "common.cuh"

namespace hp {

constexpr int TN_T = 128;         // points per tile
constexpr int TN_TP = TN_T + 4;   // activation row stride (floats)
constexpr int TN_THREADS = 512;
constexpr int TN_WARPS = TN_THREADS / 32;
constexpr int TN_MAX_LAYERS = 16;
constexpr int TN_MAX_GRID = 256;  // fast-path CTAs (one per SM)

struct TNArgs {
    const float *weights;   // [B][W]
    const float *points;    // [B][N][3] (pstride = 3N) or shared [N][3] (pstride = 0)
    long long pstride;
    float *out;             // forward: [B][N][3] or [B][3][N]
    const float *gout;      // backward: same layout as out
    float *gweights;        // [B][W]
    float *gpoints;         // [B][N][3] or nullptr
    float *partial;         // [grid][S][W]: per-CTA partial dW, slot = sample - first sample of the CTA's range
    unsigned int *counters; // [B], zeroed by the launcher
    int B, N, W, S;         // S = partial slots per CTA (backward)
    int channels_first;
    int offw[TN_MAX_LAYERS], offb[TN_MAX_LAYERS];  // float offsets inside a sample's weight vector; offb < 0: no bias
};

// ---- staging helpers ---------------------------------------------------------------------------------
// W[OUT][K] (flat, row-major) -> smem rows of K+4 floats (KPAD = K rounded up to 4; extra columns zeroed).
template <int K, int OUT>
__device__ __forceinline__ void tn_stage_weights(const float *__restrict__ Wg, float *__restrict__ Ws, int tid) {
    constexpr int KR = (K + 3) & ~3;
    constexpr int KP = KR + 4;
    if (K == KR) {
        for (int i = tid; i < OUT * K; i += TN_THREADS) {
            const int o = i / K, k = i - o * K;
            Ws[o * KP + k] = __ldg(Wg + i);
        }
    } else {
        for (int i = tid; i < OUT * KR; i += TN_THREADS) {
            const int o = i / KR, k = i - o * KR;
            Ws[o * KP + k] = (k < K) ? __ldg(Wg + o * K + k) : 0.f;
        }
    }
}
// 3-row matrices (the output layer) are staged as 4 rows, the 4th all zero.
template <int K>
__device__ __forceinline__ void tn_stage_weights_out3(const float *__restrict__ Wg, float *__restrict__ Ws, int tid) {
    constexpr int KP = K + 4;
    for (int i = tid; i < 4 * K; i += TN_THREADS) {
        const int o = i / K, k = i - o * K;
        Ws[o * KP + k] = (o < 3) ? __ldg(Wg + i) : 0.f;
    }
}

// points tile -> X[4][TP] (row 3 zero, out-of-range points zero)
__device__ __forceinline__ void tn_load_xyz_tile(const float *__restrict__ src /* [N][3] of this sample */, int n0, int N,
                                                 float *__restrict__ X, int tid) {
    for (int i = tid; i < TN_T * 3; i += TN_THREADS) {
        const int p = i / 3, c = i - p * 3;
        X[c * TN_TP + p] = (n0 + p < N) ? __ldg(src + (size_t)(n0 + p) * 3 + c) : 0.f;
    }
    for (int p = tid; p < TN_T; p += TN_THREADS) X[3 * TN_TP + p] = 0.f;
}
// same for a [3][N] (channels-first) source
__device__ __forceinline__ void tn_load_cf_tile(const float *__restrict__ src /* [3][N] */, int n0, int N,
                                                float *__restrict__ X, int tid) {
    for (int i = tid; i < TN_T * 3; i += TN_THREADS) {
        const int c = i / TN_T, p = i - c * TN_T;
        X[c * TN_TP + p] = (n0 + p < N) ? __ldg(src + (size_t)c * N + n0 + p) : 0.f;
    }
    for (int p = tid; p < TN_T; p += TN_THREADS) X[3 * TN_TP + p] = 0.f;
}

// ---- forward layer: out[o][p] = act(b[o] + sum_k W[o][k] in[k][p]) --------------------------------------
// lane = (og, pg): pg = lane & 7 -> 4 points, og = lane >> 3 -> channels og + 4j.  Warp tile 32 points x 4*OT channels.
template <int K, int OUT, int OT, bool RELU>
__device__ __forceinline__ void tn_layer_fwd(const float *__restrict__ in, float *__restrict__ out,
                                             const float *__restrict__ Ws, const float *__restrict__ bs, int warp, int lane) {
    constexpr int KR = (K + 3) & ~3;
    constexpr int KP = KR + 4;
    static_assert(OUT % (4 * OT) == 0, "channel tiling");
    constexpr int NWO = OUT / (4 * OT);
    const int pg = lane & 7, og = lane >> 3;
    for (int wt = warp; wt < 4 * NWO; wt += TN_WARPS) {
        const int wp = wt & 3, wo = wt >> 2;
        const int p = wp * 32 + pg * 4;
        const int ob = wo * (4 * OT) + og;
        // accumulators as fp32x2 pairs over points (p0,p1), (p2,p3): FFMA2 with the weight as the broadcast scalar operand
        // (SASS: FFMA2 Rd, Rw.F32, Ra.F32x2, Rd) -- half the issue slots of scalar FFMA, bit-identical lanes
        f32x2 acc[OT][2];
#pragma unroll
        for (int j = 0; j < OT; ++j) {
            const float b = bs[ob + 4 * j];
            acc[j][0] = acc[j][1] = pack2(b, b);
        }
#pragma unroll 2
        for (int k0 = 0; k0 < KR; k0 += 4) {
            ulonglong2 a[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) a[kk] = *reinterpret_cast<const ulonglong2 *>(in + (k0 + kk) * TN_TP + p);
#pragma unroll
            for (int j = 0; j < OT; ++j) {
                const float4 w = *reinterpret_cast<const float4 *>(Ws + (ob + 4 * j) * KP + k0);
                const f32x2 w0 = pack2(w.x, w.x), w1 = pack2(w.y, w.y), w2 = pack2(w.z, w.z), w3 = pack2(w.w, w.w);
                acc[j][0] = fma2(w0, a[0].x, acc[j][0]), acc[j][1] = fma2(w0, a[0].y, acc[j][1]);
                acc[j][0] = fma2(w1, a[1].x, acc[j][0]), acc[j][1] = fma2(w1, a[1].y, acc[j][1]);
                acc[j][0] = fma2(w2, a[2].x, acc[j][0]), acc[j][1] = fma2(w2, a[2].y, acc[j][1]);
                acc[j][0] = fma2(w3, a[3].x, acc[j][0]), acc[j][1] = fma2(w3, a[3].y, acc[j][1]);
            }
        }
#pragma unroll
        for (int j = 0; j < OT; ++j) {
            float4 r;
            unpack2(acc[j][0], r.x, r.y);
            unpack2(acc[j][1], r.z, r.w);
            if (RELU) r.x = fmaxf(r.x, 0.f), r.y = fmaxf(r.y, 0.f), r.z = fmaxf(r.z, 0.f), r.w = fmaxf(r.w, 0.f);
            *reinterpret_cast<float4 *>(out + (ob + 4 * j) * TN_TP + p) = r;
        }
    }
}

// ---- dgrad: in[k][p] <- (in[k][p] > 0 ? 1 : 0) * sum_o Z[o][p] W[o][k]   (in place; MASK=false: plain store) ----
// lane = (kg, pg): pg = lane & 7 -> 4 points, kg = lane >> 3 -> TK input channels.  Warp tile 32 points x 4*TK channels.
template <int KW, int OW, int TK, bool MASK>
__device__ __forceinline__ void tn_layer_dgrad(const float *__restrict__ Z, float *__restrict__ in,
                                               const float *__restrict__ Ws, int warp, int lane) {
    constexpr int KR = (KW + 3) & ~3;
    constexpr int KP = KR + 4;
    constexpr int V = TK >= 4 ? 4 : TK;  // channels per shared-memory vector load
    constexpr int H = TK / V;
    static_assert(KR % (4 * TK) == 0 && (V == 4 || V == 2), "channel tiling");
    constexpr int NWK = KR / (4 * TK);
    const int pg = lane & 7, kg = lane >> 3;
    for (int wt = warp; wt < 4 * NWK; wt += TN_WARPS) {
        const int wp = wt & 3, wk = wt >> 2;
        const int p = wp * 32 + pg * 4;
        const int kb = wk * (4 * TK) + kg * V;  // channel of (h, i): kb + h*4*V + i
        f32x2 acc[H][V][2];  // pairs over points, see tn_layer_fwd
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
            for (int i = 0; i < V; ++i) acc[h][i][0] = acc[h][i][1] = pack2(0.f, 0.f);
#pragma unroll 4
        for (int o = 0; o < OW; ++o) {
            const ulonglong2 z = *reinterpret_cast<const ulonglong2 *>(Z + o * TN_TP + p);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float w[4];
                if (V == 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(Ws + o * KP + kb + h * 16);
                    w[0] = t.x, w[1] = t.y, w[2] = t.z, w[3] = t.w;
                } else {
                    const float2 t = *reinterpret_cast<const float2 *>(Ws + o * KP + kb + h * 8);
                    w[0] = t.x, w[1] = t.y;
                }
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const f32x2 wi = pack2(w[i], w[i]);
                    acc[h][i][0] = fma2(wi, z.x, acc[h][i][0]), acc[h][i][1] = fma2(wi, z.y, acc[h][i][1]);
                }
            }
        }
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const int k = kb + h * 4 * V + i;
                float4 *dst = reinterpret_cast<float4 *>(in + k * TN_TP + p);
                float4 r;
                unpack2(acc[h][i][0], r.x, r.y);
                unpack2(acc[h][i][1], r.z, r.w);
                if (MASK) {
                    const float4 a = *dst;
                    r.x = a.x > 0.f ? r.x : 0.f, r.y = a.y > 0.f ? r.y : 0.f;
                    r.z = a.z > 0.f ? r.z : 0.f, r.w = a.w > 0.f ? r.w : 0.f;
                }
                *dst = r;
            }
    }
}

// ---- wgrad: acc[j][i] += sum_p Z[o_j][p] A[k_i][p];  exactly one warp tile per warp => persistent registers ----
// lane = (kg, og): og = lane & 7 -> rows og + 8j of Z, kg = lane >> 3 -> rows kg + 4i of A.
template <int OW, int KW, int TO, int TK>
struct TNWgradMap {
    static constexpr int NWO = OW / (8 * TO), NWK = KW / (4 * TK);
    static_assert(NWO * NWK == TN_WARPS && OW % (8 * TO) == 0 && KW % (4 * TK) == 0, "one warp tile per warp");
    __device__ static __forceinline__ int o(int warp, int lane, int j) { return (warp % NWO) * 8 * TO + (lane & 7) + 8 * j; }
    __device__ static __forceinline__ int k(int warp, int lane, int i) { return (warp / NWO) * 4 * TK + (lane >> 3) + 4 * i; }
};
template <int OW, int KW, int TO, int TK>
__device__ __forceinline__ void tn_layer_wgrad(const float *__restrict__ Z, const float *__restrict__ A,
                                               float (&acc)[TO * TK], int warp, int lane) {
    using M = TNWgradMap<OW, KW, TO, TK>;
    const float *zr = Z + M::o(warp, lane, 0) * TN_TP;
    const float *ar = A + M::k(warp, lane, 0) * TN_TP;
    // per tile: even / odd point partial sums as fp32x2 pairs (FFMA2: half the issue slots), folded into the
    // persistent scalar accumulators once per tile
    f32x2 t2[TO * TK];
#pragma unroll
    for (int i = 0; i < TO * TK; ++i) t2[i] = pack2(0.f, 0.f);
#pragma unroll 2
    for (int p0 = 0; p0 < TN_T; p0 += 4) {
        ulonglong2 z[TO], a[TK];
#pragma unroll
        for (int j = 0; j < TO; ++j) z[j] = *reinterpret_cast<const ulonglong2 *>(zr + j * 8 * TN_TP + p0);
#pragma unroll
        for (int i = 0; i < TK; ++i) a[i] = *reinterpret_cast<const ulonglong2 *>(ar + i * 4 * TN_TP + p0);
#pragma unroll
        for (int j = 0; j < TO; ++j)
#pragma unroll
            for (int i = 0; i < TK; ++i) {
                t2[j * TK + i] = fma2(z[j].x, a[i].x, t2[j * TK + i]);
                t2[j * TK + i] = fma2(z[j].y, a[i].y, t2[j * TK + i]);
            }
    }
#pragma unroll
    for (int i = 0; i < TO * TK; ++i) {
        float lo, hi;
        unpack2(t2[i], lo, hi);
        acc[i] += lo + hi;
    }
}
template <int OW, int KW, int TO, int TK>
__device__ __forceinline__ void tn_store_wgrad(const float (&acc)[TO * TK], float *__restrict__ dst /* [OW][KW] */, int warp, int lane) {
    using M = TNWgradMap<OW, KW, TO, TK>;
#pragma unroll
    for (int j = 0; j < TO; ++j)
#pragma unroll
        for (int i = 0; i < TK; ++i) dst[M::o(warp, lane, j) * KW + M::k(warp, lane, i)] = acc[j * TK + i];
}

// row sums / row dots over the 128 points of a tile (the small gradients: biases, dW of the 3-wide layers)
__device__ __forceinline__ float tn_row_sum(const float *__restrict__ z) {
    float s = 0.f;
#pragma unroll 8
    for (int p = 0; p < TN_T; p += 4) {
        const float4 v = *reinterpret_cast<const float4 *>(z + p);
        s += (v.x + v.y) + (v.z + v.w);
    }
    return s;
}
__device__ __forceinline__ float tn_row_dot(const float *__restrict__ z, const float *__restrict__ a) {
    float s = 0.f;
#pragma unroll 8
    for (int p = 0; p < TN_T; p += 4) {
        const float4 u = *reinterpret_cast<const float4 *>(z + p), v = *reinterpret_cast<const float4 *>(a + p);
        s = __fmaf_rn(u.x, v.x, s), s = __fmaf_rn(u.y, v.y, s), s = __fmaf_rn(u.z, v.z, s), s = __fmaf_rn(u.w, v.w, s);
    }
    return s;
}

// Fold the per-CTA partial gradients of a sample that several CTAs share: g[i] = sum over the contributing slots in ascending CTA
// order (deterministic).  Eight elements per thread in flight: the loop is bound by L2 latency, not bandwidth.
template <int NT>
__device__ __forceinline__ void tn_fold_partials(const float *__restrict__ partial, const unsigned int *__restrict__ slot_of, int ncontrib,
                                                 int W, float *__restrict__ g, int tid) {
    for (int i0 = tid; i0 < W; i0 += 8 * NT) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = 0.f;
        for (int q = 0; q < ncontrib; ++q) {
            const float *src = partial + (size_t)slot_of[q] * W;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (i0 + u * NT < W) v[u] += __ldcg(src + i0 + u * NT);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (i0 + u * NT < W) g[i0 + u * NT] = v[u];
    }
}

// ---- fast path: 3 -> 32 -> 64 -> 128 -> 64 -> 3 -----------------------------------------------------------
constexpr int C1 = 32, C2 = 64, C3 = 128, C4 = 64;
// forward shared-memory map (floats)
constexpr int TNF_BUFA = 0;                                  // [128][TP]   A1 / A3
constexpr int TNF_BUFB = TNF_BUFA + C3 * TN_TP;              // [64][TP]    X / A2 / A4
constexpr int TNF_W1 = TNF_BUFB + C2 * TN_TP;                // [32][8]
constexpr int TNF_W2 = TNF_W1 + C1 * 8;                      // [64][36]
constexpr int TNF_W3 = TNF_W2 + C2 * (C1 + 4);               // [128][68]
constexpr int TNF_W4 = TNF_W3 + C3 * (C2 + 4);               // [64][132]
constexpr int TNF_W5 = TNF_W4 + C4 * (C3 + 4);               // [4][68]
constexpr int TNF_BIAS = TNF_W5 + 4 * (C4 + 4);              // 32 + 64 + 128 + 64 + 4
constexpr int TNF_FLOATS = TNF_BIAS + C1 + C2 + C3 + C4 + 4;
constexpr size_t TNF_SMEM = (size_t)TNF_FLOATS * sizeof(float);

__device__ __forceinline__ void tn_stage_biases(const TNArgs &a, const float *__restrict__ wg, float *__restrict__ bs, int tid) {
    // bs: b1[32] b2[64] b3[128] b4[64] b5[4]
    for (int i = tid; i < C1 + C2 + C3 + C4 + 4; i += TN_THREADS) {
        int l, o;
        if (i < C1) l = 0, o = i;
        else if (i < C1 + C2) l = 1, o = i - C1;
        else if (i < C1 + C2 + C3) l = 2, o = i - C1 - C2;
        else if (i < C1 + C2 + C3 + C4) l = 3, o = i - C1 - C2 - C3;
        else l = 4, o = i - C1 - C2 - C3 - C4;
        bs[i] = (a.offb[l] >= 0 && !(l == 4 && o >= 3)) ? __ldg(wg + a.offb[l] + o) : 0.f;
    }
}

__device__ __forceinline__ void tn_store_out_tile(const TNArgs &a, int b, int n0, const float *__restrict__ Y /* [4][TP] */, int tid) {
    const int N = a.N;
    if (a.channels_first) {
        float *dst = a.out + (size_t)b * 3 * N;
        for (int i = tid; i < 3 * TN_T; i += TN_THREADS) {
            const int c = i / TN_T, p = i - c * TN_T;
            if (n0 + p < N) dst[(size_t)c * N + n0 + p] = Y[c * TN_TP + p];
        }
    } else {
        float *dst = a.out + ((size_t)b * N + n0) * 3;
        const int lim = min(TN_T, N - n0) * 3;
        for (int i = tid; i < lim; i += TN_THREADS) {
            const int p = i / 3, c = i - p * 3;
            dst[i] = Y[c * TN_TP + p];
        }
    }
}

__global__ void __launch_bounds__(TN_THREADS, 1) tn_forward_kernel(const TNArgs a) {
    extern __shared__ __align__(16) float sm[];
    float *bufA = sm + TNF_BUFA, *bufB = sm + TNF_BUFB;
    float *W1 = sm + TNF_W1, *W2 = sm + TNF_W2, *W3 = sm + TNF_W3, *W4 = sm + TNF_W4, *W5 = sm + TNF_W5, *bs = sm + TNF_BIAS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // flat (sample, tile) list split evenly over the CTAs: every SM gets the same number of tiles (+-1)
    const int ntiles = (a.N + TN_T - 1) / TN_T;
    const long long TT = (long long)a.B * ntiles;
    const long long f0 = (long long)blockIdx.x * TT / gridDim.x, f1 = (long long)(blockIdx.x + 1) * TT / gridDim.x;
    int cur = -1;
    for (long long f = f0; f < f1; ++f) {
        const int b = (int)(f / ntiles), t = (int)(f - (long long)b * ntiles);
        const float *pts = a.points + (size_t)b * a.pstride;
        if (b != cur) {  // (re)stage this sample's weights; every reader of the old ones is past the last barrier
            const float *wg = a.weights + (size_t)b * a.W;
            tn_stage_weights<3, C1>(wg + a.offw[0], W1, tid);
            tn_stage_weights<C1, C2>(wg + a.offw[1], W2, tid);
            tn_stage_weights<C2, C3>(wg + a.offw[2], W3, tid);
            tn_stage_weights<C3, C4>(wg + a.offw[3], W4, tid);
            tn_stage_weights_out3<C4>(wg + a.offw[4], W5, tid);
            tn_stage_biases(a, wg, bs, tid);
            cur = b;
        }
        const int n0 = t * TN_T;
        __syncthreads();  // previous tile's Y consumed; weights visible
        tn_load_xyz_tile(pts, n0, a.N, bufB, tid);
        __syncthreads();
        tn_layer_fwd<3, C1, 2, true>(bufB, bufA, W1, bs, warp, lane);
        __syncthreads();
        tn_layer_fwd<C1, C2, 4, true>(bufA, bufB, W2, bs + C1, warp, lane);
        __syncthreads();
        tn_layer_fwd<C2, C3, 8, true>(bufB, bufA, W3, bs + C1 + C2, warp, lane);
        __syncthreads();
        tn_layer_fwd<C3, C4, 4, true>(bufA, bufB, W4, bs + C1 + C2 + C3, warp, lane);
        __syncthreads();
        tn_layer_fwd<C4, 4, 1, false>(bufB, bufA, W5, bs + C1 + C2 + C3 + C4, warp, lane);
        __syncthreads();
        tn_store_out_tile(a, b, n0, bufA, tid);
    }
}

// backward shared-memory map (floats)
constexpr int TNB_A1 = 0;                           // [32][TP]
constexpr int TNB_A2 = TNB_A1 + C1 * TN_TP;         // [64][TP]
constexpr int TNB_A3 = TNB_A2 + C2 * TN_TP;         // [128][TP]
constexpr int TNB_A4 = TNB_A3 + C3 * TN_TP;         // [64][TP]
constexpr int TNB_X = TNB_A4 + C4 * TN_TP;          // [4][TP]
constexpr int TNB_G = TNB_X + 4 * TN_TP;            // [4][TP]  dY (row 3 zero)
constexpr int TNB_WBUF = TNB_G + 4 * TN_TP;         // 2 x [128*68]
constexpr int TNB_WSZ = C3 * (C2 + 4);              // 8704 >= 64*132 = 8448
constexpr int TNB_BIAS = TNB_WBUF + 2 * TNB_WSZ;
constexpr int TNB_FLOATS = TNB_BIAS + C1 + C2 + C3 + C4 + 4;
constexpr size_t TNB_SMEM = (size_t)TNB_FLOATS * sizeof(float);
static_assert(TNB_SMEM <= 227 * 1024, "backward tile does not fit in shared memory");
static_assert(C4 * (C3 + 4) <= TNB_WSZ, "weight buffer");

template <bool GRAD_POINTS>
__global__ void __launch_bounds__(TN_THREADS, 1) tn_backward_kernel(const TNArgs a) {
    extern __shared__ __align__(16) float sm[];
    float *A1 = sm + TNB_A1, *A2 = sm + TNB_A2, *A3 = sm + TNB_A3, *A4 = sm + TNB_A4, *X = sm + TNB_X, *G5 = sm + TNB_G;
    float *wb0 = sm + TNB_WBUF, *wb1 = wb0 + TNB_WSZ, *bs = sm + TNB_BIAS;
    __shared__ int is_last;
    __shared__ unsigned int slot_of[TN_MAX_GRID];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // flat (sample, tile) list split evenly over the CTAs (every SM busy); a CTA's range may straddle samples
    const int ntiles = (a.N + TN_T - 1) / TN_T;
    const long long TT = (long long)a.B * ntiles;
    const long long G = gridDim.x;
    const long long f0 = (long long)blockIdx.x * TT / G, f1 = (long long)(blockIdx.x + 1) * TT / G;
    const int first_sample = (int)(f0 / ntiles);

    float dW4[16], dW3[16], dW2[4];
    // small gradients, one role per warp: small0 = dW5 (warps 0-5) | db5 (6) | db4 (7-8) | db3 (9-12) | db2 (13-14) | db1 (15)
    //                                     small1 = dW1 (warps 0-2)
    float small0, small1;
    auto reset_acc = [&]() {
#pragma unroll
        for (int i = 0; i < 16; ++i) dW4[i] = 0.f, dW3[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) dW2[i] = 0.f;
        small0 = 0.f, small1 = 0.f;
    };
    // Write this CTA's accumulated dW of sample `b` (flat weight order).  If other CTAs also hold tiles of `b`, the
    // vector goes to this CTA's partial slot and the last CTA to arrive folds all slots in ascending CTA order.
    auto flush = [&](int b) {
        const long long s0 = (long long)b * ntiles, s1 = s0 + ntiles - 1;  // flat tiles of sample b
        const int c_lo = (int)(((s0 + 1) * G - 1) / TT), c_hi = (int)(((s1 + 1) * G - 1) / TT);
        const bool shared_sample = c_lo != c_hi;
        float *dst = shared_sample ? a.partial + ((size_t)blockIdx.x * a.S + (b - first_sample)) * a.W : a.gweights + (size_t)b * a.W;
        tn_store_wgrad<C4, C3, 4, 4>(dW4, dst + a.offw[3], warp, lane);
        tn_store_wgrad<C3, C2, 4, 4>(dW3, dst + a.offw[2], warp, lane);
        tn_store_wgrad<C2, C1, 2, 2>(dW2, dst + a.offw[1], warp, lane);
        if (warp < 6) dst[a.offw[4] + tid] = small0;  // dW5[o][k] at o*64 + k = tid
        else if (warp == 6) { if (lane < 3 && a.offb[4] >= 0) dst[a.offb[4] + lane] = small0; }
        else if (warp <= 8) { if (a.offb[3] >= 0) dst[a.offb[3] + (warp - 7) * 32 + lane] = small0; }
        else if (warp <= 12) { if (a.offb[2] >= 0) dst[a.offb[2] + (warp - 9) * 32 + lane] = small0; }
        else if (warp <= 14) { if (a.offb[1] >= 0) dst[a.offb[1] + (warp - 13) * 32 + lane] = small0; }
        else { if (a.offb[0] >= 0) dst[a.offb[0] + lane] = small0; }
        if (warp < 3) dst[a.offw[0] + tid] = small1;  // dW1[o][c] at o*3 + c = tid
        if (shared_sample) {
            __threadfence();
            __syncthreads();
            if (tid == 0) is_last = (atomicAdd(a.counters + b, 1u) == (unsigned)(c_hi - c_lo));
            __syncthreads();
            if (is_last) {
                __threadfence();
                // slot of sample b in every contributing CTA's partial block (64-bit divisions hoisted out of the fold)
                const int ncontrib = c_hi - c_lo + 1;
                for (int q = tid; q < ncontrib; q += TN_THREADS) {
                    const int c = c_lo + q;
                    const int fs = (int)(((long long)c * TT / G) / ntiles);
                    slot_of[q] = (unsigned)(c * a.S + (b - fs));
                }
                __syncthreads();
                float *g = a.gweights + (size_t)b * a.W;
                tn_fold_partials<TN_THREADS>(a.partial, slot_of, ncontrib, a.W, g, tid);
            }
            __syncthreads();  // is_last reusable
        }
    };

    reset_acc();
    int cur = -1;
    for (long long f = f0; f < f1; ++f) {
        const int b = (int)(f / ntiles), t = (int)(f - (long long)b * ntiles);
        if (b != cur) {
            if (cur >= 0) {
                flush(cur);
                reset_acc();
            }
            __syncthreads();  // every reader of the old biases is done
            tn_stage_biases(a, a.weights + (size_t)b * a.W, bs, tid);
            cur = b;
        }
        const float *wg = a.weights + (size_t)b * a.W;
        const float *pts = a.points + (size_t)b * a.pstride;
        const float *gy = a.gout + (size_t)b * a.N * 3;
        const int n0 = t * TN_T;
        __syncthreads();  // previous tile fully consumed
        tn_load_xyz_tile(pts, n0, a.N, X, tid);
        if (a.channels_first) tn_load_cf_tile(gy, n0, a.N, G5, tid);
        else tn_load_xyz_tile(gy, n0, a.N, G5, tid);
        tn_stage_weights<3, C1>(wg + a.offw[0], wb0, tid);
        __syncthreads();
        // ---- recompute the forward of this tile ----
        tn_stage_weights<C1, C2>(wg + a.offw[1], wb1, tid);
        tn_layer_fwd<3, C1, 2, true>(X, A1, wb0, bs, warp, lane);
        __syncthreads();
        tn_stage_weights<C2, C3>(wg + a.offw[2], wb0, tid);
        tn_layer_fwd<C1, C2, 4, true>(A1, A2, wb1, bs + C1, warp, lane);
        __syncthreads();
        tn_stage_weights<C3, C4>(wg + a.offw[3], wb1, tid);
        tn_layer_fwd<C2, C3, 8, true>(A2, A3, wb0, bs + C1 + C2, warp, lane);
        __syncthreads();
        tn_stage_weights_out3<C4>(wg + a.offw[4], wb0, tid);
        tn_layer_fwd<C3, C4, 4, true>(A3, A4, wb1, bs + C1 + C2 + C3, warp, lane);
        __syncthreads();
        // ---- layer 5: Z5 = dY ----
        if (warp < 6) small0 += tn_row_dot(G5 + (tid >> 6) * TN_TP, A4 + (tid & 63) * TN_TP);       // dW5[o][k], o = tid/64
        else if (warp == 6) { if (lane < 3) small0 += tn_row_sum(G5 + lane * TN_TP); }                // db5
        __syncthreads();
        tn_layer_dgrad<C4, 3, 4, true>(G5, A4, wb0, warp, lane);  // A4 <- Z4     (wb1 still holds W4)
        __syncthreads();
        // ---- layer 4 ----
        tn_stage_weights<C2, C3>(wg + a.offw[2], wb0, tid);
        tn_layer_wgrad<C4, C3, 4, 4>(A4, A3, dW4, warp, lane);
        if (warp == 7 || warp == 8) small0 += tn_row_sum(A4 + ((warp - 7) * 32 + lane) * TN_TP);     // db4
        __syncthreads();
        tn_layer_dgrad<C3, C4, 8, true>(A4, A3, wb1, warp, lane);  // A3 <- Z3
        __syncthreads();
        // ---- layer 3 ----
        tn_stage_weights<C1, C2>(wg + a.offw[1], wb1, tid);
        tn_layer_wgrad<C3, C2, 4, 4>(A3, A2, dW3, warp, lane);
        if (warp >= 9 && warp <= 12) small0 += tn_row_sum(A3 + ((warp - 9) * 32 + lane) * TN_TP);    // db3
        __syncthreads();
        tn_layer_dgrad<C2, C3, 4, true>(A3, A2, wb0, warp, lane);  // A2 <- Z2
        __syncthreads();
        // ---- layer 2 ----
        if (GRAD_POINTS) tn_stage_weights<3, C1>(wg + a.offw[0], wb0, tid);
        tn_layer_wgrad<C2, C1, 2, 2>(A2, A1, dW2, warp, lane);
        if (warp == 13 || warp == 14) small0 += tn_row_sum(A2 + ((warp - 13) * 32 + lane) * TN_TP);  // db2
        __syncthreads();
        tn_layer_dgrad<C1, C2, 2, true>(A2, A1, wb1, warp, lane);  // A1 <- Z1
        __syncthreads();
        // ---- layer 1 ----
        if (warp < 3) small1 += tn_row_dot(A1 + (tid / 3) * TN_TP, X + (tid % 3) * TN_TP);            // dW1[o][c], o = tid/3
        if (warp == 15) small0 += tn_row_sum(A1 + lane * TN_TP);                                      // db1
        if (GRAD_POINTS) {
            // dX[c][p] = sum_o Z1[o][p] W1[o][c]
            if (tid < 3 * TN_T) {
                const int c = tid / TN_T, p = tid - c * TN_T;
                float sx = 0.f;
#pragma unroll 8
                for (int o = 0; o < C1; ++o) sx = __fmaf_rn(A1[o * TN_TP + p], wb0[o * 8 + c], sx);
                if (n0 + p < a.N) a.gpoints[((size_t)b * a.N + n0 + p) * 3 + c] = sx;
            }
        }
    }
    if (cur >= 0) flush(cur);
}

}  // namespace hp
#include "target_network_mma.cuh"
#include "target_network_tc5.cuh"
namespace hp {

// ---- generic path: any widths ------------------------------------------------------------------------
constexpr int TNG_T = 32;        // points per tile
constexpr int TNG_THREADS = 256;
struct TNGenArgs {
    TNArgs base;
    int n_layers;
    int dims[TN_MAX_LAYERS + 1];
    int act_off[TN_MAX_LAYERS + 1];  // float offset of layer l's activation rows [dims[l]][TNG_T] in shared memory
    int sum_dims;
    int max_dim;
};

__global__ void __launch_bounds__(TNG_THREADS) tn_generic_forward_kernel(const TNGenArgs g) {
    extern __shared__ __align__(16) float sm[];  // two ping-pong buffers [max_dim][TNG_T]
    const TNArgs &a = g.base;
    const int tid = threadIdx.x;
    const int b = blockIdx.y, n0 = blockIdx.x * TNG_T;
    const float *wg = a.weights + (size_t)b * a.W;
    const float *pts = a.points + (size_t)b * a.pstride;
    float *cur = sm, *nxt = sm + (size_t)g.max_dim * TNG_T;
    for (int i = tid; i < 3 * TNG_T; i += TNG_THREADS) {
        const int p = i / 3, c = i - p * 3;
        cur[c * TNG_T + p] = (n0 + p < a.N) ? __ldg(pts + (size_t)(n0 + p) * 3 + c) : 0.f;
    }
    __syncthreads();
    for (int l = 0; l < g.n_layers; ++l) {
        const int K = g.dims[l], O = g.dims[l + 1];
        const float *W = wg + a.offw[l];
        const bool last = (l + 1 == g.n_layers);
        for (int i = tid; i < O * TNG_T; i += TNG_THREADS) {
            const int o = i / TNG_T, p = i - o * TNG_T;
            float acc = 0.f;
            for (int k = 0; k < K; ++k) acc = __fmaf_rn(cur[k * TNG_T + p], __ldg(W + (size_t)o * K + k), acc);
            if (a.offb[l] >= 0) acc += __ldg(wg + a.offb[l] + o);
            if (!last) acc = fmaxf(acc, 0.f);
            nxt[o * TNG_T + p] = acc;
        }
        __syncthreads();
        float *t = cur; cur = nxt; nxt = t;
    }
    for (int i = tid; i < 3 * TNG_T; i += TNG_THREADS) {
        const int c = i / TNG_T, p = i - c * TNG_T;
        if (n0 + p < a.N) {
            if (a.channels_first) a.out[((size_t)b * 3 + c) * a.N + n0 + p] = cur[c * TNG_T + p];
            else a.out[((size_t)b * a.N + n0 + p) * 3 + c] = cur[c * TNG_T + p];
        }
    }
}

// One CTA per sample walks all tiles: every gradient element is owned by one thread and accumulated in
// global memory in tile order => deterministic.
__global__ void __launch_bounds__(TNG_THREADS) tn_generic_backward_kernel(const TNGenArgs g) {
    extern __shared__ __align__(16) float sm[];  // acts [sum_dims][T] | Z ping-pong 2 x [max_dim][T]
    const TNArgs &a = g.base;
    const int tid = threadIdx.x, b = blockIdx.x, L = g.n_layers;
    const float *wg = a.weights + (size_t)b * a.W;
    const float *pts = a.points + (size_t)b * a.pstride;
    float *gw = a.gweights + (size_t)b * a.W;
    float *acts = sm;
    float *z0 = sm + (size_t)g.sum_dims * TNG_T, *z1 = z0 + (size_t)g.max_dim * TNG_T;
    const int ntiles = (a.N + TNG_T - 1) / TNG_T;
    for (int t = 0; t < ntiles; ++t) {
        const int n0 = t * TNG_T;
        __syncthreads();
        for (int i = tid; i < 3 * TNG_T; i += TNG_THREADS) {
            const int p = i / 3, c = i - p * 3;
            acts[c * TNG_T + p] = (n0 + p < a.N) ? __ldg(pts + (size_t)(n0 + p) * 3 + c) : 0.f;
            float gv = 0.f;
            if (n0 + p < a.N)
                gv = a.channels_first ? __ldg(a.gout + ((size_t)b * 3 + c) * a.N + n0 + p)
                                      : __ldg(a.gout + ((size_t)b * a.N + n0 + p) * 3 + c);
            z0[c * TNG_T + p] = gv;
        }
        __syncthreads();
        for (int l = 0; l + 1 < L; ++l) {  // hidden activations (the output itself is not needed)
            const int K = g.dims[l], O = g.dims[l + 1];
            const float *W = wg + a.offw[l];
            const float *cur = acts + g.act_off[l];
            float *nxt = acts + g.act_off[l + 1];
            for (int i = tid; i < O * TNG_T; i += TNG_THREADS) {
                const int o = i / TNG_T, p = i - o * TNG_T;
                float acc = 0.f;
                for (int k = 0; k < K; ++k) acc = __fmaf_rn(cur[k * TNG_T + p], __ldg(W + (size_t)o * K + k), acc);
                if (a.offb[l] >= 0) acc += __ldg(wg + a.offb[l] + o);
                nxt[o * TNG_T + p] = fmaxf(acc, 0.f);
            }
            __syncthreads();
        }
        float *Z = z0, *Zn = z1;
        for (int l = L - 1; l >= 0; --l) {
            const int K = g.dims[l], O = g.dims[l + 1];
            const float *W = wg + a.offw[l];
            const float *A = acts + g.act_off[l];
            for (int i = tid; i < O * K; i += TNG_THREADS) {  // dW[o][k]
                const int o = i / K, k = i - o * K;
                float s = 0.f;
                for (int p = 0; p < TNG_T; ++p) s = __fmaf_rn(Z[o * TNG_T + p], A[k * TNG_T + p], s);
                gw[a.offw[l] + i] = (t == 0) ? s : gw[a.offw[l] + i] + s;
            }
            if (a.offb[l] >= 0)
                for (int o = tid; o < O; o += TNG_THREADS) {
                    float s = 0.f;
                    for (int p = 0; p < TNG_T; ++p) s += Z[o * TNG_T + p];
                    gw[a.offb[l] + o] = (t == 0) ? s : gw[a.offb[l] + o] + s;
                }
            if (l > 0 || a.gpoints != nullptr) {
                for (int i = tid; i < K * TNG_T; i += TNG_THREADS) {  // dA[k][p], ReLU mask of layer l's input
                    const int k = i / TNG_T, p = i - k * TNG_T;
                    float s = 0.f;
                    for (int o = 0; o < O; ++o) s = __fmaf_rn(Z[o * TNG_T + p], __ldg(W + (size_t)o * K + k), s);
                    if (l > 0) s = A[k * TNG_T + p] > 0.f ? s : 0.f;
                    Zn[k * TNG_T + p] = s;
                }
            }
            __syncthreads();
            float *tmp = Z; Z = Zn; Zn = tmp;
        }
        if (a.gpoints != nullptr) {
            for (int i = tid; i < 3 * TNG_T; i += TNG_THREADS) {
                const int p = i / 3, c = i - p * 3;
                if (n0 + p < a.N) a.gpoints[((size_t)b * a.N + n0 + p) * 3 + c] = Z[c * TNG_T + p];
            }
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------
// 0 (default): 3xTF32 on the tensor cores -- forward on tcgen05 (target_network_tc5.cuh), backward on mma.sync (target_network_mma.cuh);
// 1: the FP32-pipe kernels above;  2: as 0 with the forward on mma.sync as well
static int g_tn_mode = 0;

static bool tn_is_fast_shape(int n_layers, const int *dims) {
    return n_layers == 5 && dims[0] == 3 && dims[1] == C1 && dims[2] == C2 && dims[3] == C3 && dims[4] == C4 && dims[5] == 3;
}

static int tn_fill_offsets(TNArgs &a, int n_layers, const int *dims, int use_bias, const char *who) {
    HP_REQUIRE(n_layers >= 1 && n_layers <= TN_MAX_LAYERS, "%s: n_layers=%d outside [1,%d]", who, n_layers, TN_MAX_LAYERS);
    HP_REQUIRE(dims != nullptr, "%s: dims is null", who);
    HP_REQUIRE(dims[0] == 3 && dims[n_layers] == 3, "%s: the network must map 3 -> 3 coordinates (dims[0]=%d, dims[last]=%d)", who,
               dims[0], dims[n_layers]);
    long long off = 0;
    for (int l = 0; l < n_layers; ++l) {
        HP_REQUIRE(dims[l] > 0 && dims[l + 1] > 0, "%s: non-positive layer width", who);
        a.offw[l] = (int)off;
        off += (long long)dims[l] * dims[l + 1];
        a.offb[l] = use_bias ? (int)off : -1;
        if (use_bias) off += dims[l + 1];
        HP_REQUIRE(off < (1LL << 30), "%s: weight vector too long", who);
    }
    a.W = (int)off;
    return HP_OK;
}

// fast-path launch geometry: one CTA per SM (never more CTAs than tiles); slots = most samples one CTA's range can touch
static void tn_geometry(int b, int n, int &grid, int &slots) {
    const long long ntiles = (n + TN_T - 1) / TN_T, TT = (long long)b * ntiles;
    const long long sms = sm_count();
    grid = (int)(TT < sms ? TT : sms);
    if (grid > TN_MAX_GRID) grid = TN_MAX_GRID;
    const long long per = (TT + grid - 1) / grid;  // tiles per CTA, upper bound
    slots = (int)((per + ntiles - 2) / ntiles + 1);
}

static int tn_generic_args(TNGenArgs &g, int n_layers, const int *dims, size_t &smem_fwd, size_t &smem_bwd) {
    g.n_layers = n_layers;
    int sum = 0, mx = 0;
    for (int l = 0; l <= n_layers; ++l) {
        g.dims[l] = dims[l];
        if (l < n_layers) {  // activations kept for backward: input + hidden layers
            g.act_off[l] = sum * TNG_T;
            sum += dims[l];
        }
        mx = dims[l] > mx ? dims[l] : mx;
    }
    g.act_off[n_layers] = 0;
    g.sum_dims = sum;
    g.max_dim = mx;
    smem_fwd = (size_t)2 * mx * TNG_T * sizeof(float);
    smem_bwd = ((size_t)sum + 2 * (size_t)mx) * TNG_T * sizeof(float);
    return HP_OK;
}

}  // namespace hp

using namespace hp;

#ifdef HP_TM_TRACE
extern "C" __attribute__((visibility("default"))) int hp_debug_tn_trace(unsigned long long *out_host) {
    return (int)cudaMemcpyFromSymbol(out_host, hp::g_tm_trace, sizeof(hp::g_tm_trace));
}
extern "C" __attribute__((visibility("default"))) int hp_debug_tn_cta(unsigned long long *out_host) {
    return (int)cudaMemcpyFromSymbol(out_host, hp::g_tm_cta, sizeof(hp::g_tm_cta));
}
#endif

extern "C" int hp_target_network_set_mode(int mode) {
    HP_REQUIRE(mode >= 0 && mode <= 2, "hp_target_network_set_mode: mode %d is neither 0 (3xTF32: tcgen05 forward, mma.sync backward) nor 1 (FP32 pipe) nor 2 (3xTF32, mma.sync only)", mode);
    g_tn_mode = mode;
    return HP_OK;
}

extern "C" long long hp_target_network_num_weights(int n_layers, const int *dims, int use_bias) {
    if (n_layers < 1 || n_layers > TN_MAX_LAYERS || dims == nullptr) return -1;
    long long off = 0;
    for (int l = 0; l < n_layers; ++l) {
        if (dims[l] <= 0 || dims[l + 1] <= 0) return -1;
        off += (long long)dims[l] * dims[l + 1] + (use_bias ? dims[l + 1] : 0);
    }
    return off;
}

extern "C" int hp_target_network_forward(int b, int n, int n_layers, const int *dims, int use_bias, const float *weights,
                                         const float *points, long long points_batch_stride, float *out,
                                         int channels_first, void *stream_v) {
    HP_REQUIRE(b >= 0 && n >= 0, "hp_target_network_forward: negative size (b=%d n=%d)", b, n);
    TNGenArgs g = {};
    TNArgs &a = g.base;
    int rc = tn_fill_offsets(a, n_layers, dims, use_bias, "hp_target_network_forward");
    if (rc != HP_OK) return rc;
    if (b == 0 || n == 0) return HP_OK;
    HP_REQUIRE(weights && points && out, "hp_target_network_forward: null pointer");
    HP_REQUIRE(points_batch_stride == 0 || points_batch_stride >= (long long)n * 3,
               "hp_target_network_forward: points_batch_stride must be 0 (shared cloud) or >= 3n");
    cudaStream_t stream = (cudaStream_t)stream_v;
    a.weights = weights, a.points = points, a.pstride = points_batch_stride, a.out = out;
    a.B = b, a.N = n, a.channels_first = channels_first ? 1 : 0;
    if (tn_is_fast_shape(n_layers, dims) && g_tn_mode == 0) {
        const long long tiles = (long long)b * ((n + TN_T - 1) / TN_T), sms = sm_count();
        const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
        static SmemAttrCache tattr;
        HP_CUDA(ensure_dynamic_smem(tn_tc5_forward_kernel, T5_SMEM, tattr));
        tn_tc5_forward_kernel<<<grid, T5_THREADS, T5_SMEM, stream>>>(a);
        HP_LAUNCH_CHECK("tn_tc5_forward_kernel");
        return HP_OK;
    }
    if (tn_is_fast_shape(n_layers, dims) && g_tn_mode == 2) {
        const long long units = (long long)b * ((n + 15) / 16), sms = sm_count();
        const unsigned grid = (unsigned)(units < sms ? units : sms);
        static SmemAttrCache mattr;
        HP_CUDA(ensure_dynamic_smem(tn_mma_forward_kernel, TMF_SMEM, mattr));
        tn_mma_forward_kernel<<<grid, TMF_THREADS, TMF_SMEM, stream>>>(a);
        HP_LAUNCH_CHECK("tn_mma_forward_kernel");
        return HP_OK;
    }
    if (tn_is_fast_shape(n_layers, dims)) {
        int grid;
        tn_geometry(b, n, grid, a.S);
        static SmemAttrCache attr;
        HP_CUDA(ensure_dynamic_smem(tn_forward_kernel, TNF_SMEM, attr));
        tn_forward_kernel<<<(unsigned)grid, TN_THREADS, TNF_SMEM, stream>>>(a);
        HP_LAUNCH_CHECK("tn_forward_kernel");
        return HP_OK;
    }
    size_t smem_f, smem_b;
    tn_generic_args(g, n_layers, dims, smem_f, smem_b);
    if (smem_f > 200 * 1024) {
        set_error("hp_target_network_forward: layer width %d needs %zu bytes of shared memory (limit 200 KB)", g.max_dim, smem_f);
        return HP_ERR_UNSUPPORTED;
    }
    HP_REQUIRE(b <= 65535, "hp_target_network_forward: batch %d > 65535 on the generic path; split the batch", b);
    static SmemAttrCache gattr;
    if (smem_f > 48 * 1024) HP_CUDA(ensure_dynamic_smem(tn_generic_forward_kernel, smem_f, gattr));
    tn_generic_forward_kernel<<<dim3((n + TNG_T - 1) / TNG_T, b), TNG_THREADS, smem_f, stream>>>(g);
    HP_LAUNCH_CHECK("tn_generic_forward_kernel");
    return HP_OK;
}

extern "C" size_t hp_target_network_backward_workspace_bytes(int b, int n, int n_layers, const int *dims, int use_bias) {
    if (b <= 0 || n <= 0 || n_layers < 1 || n_layers > TN_MAX_LAYERS || dims == nullptr) return 16;
    const long long W = hp_target_network_num_weights(n_layers, dims, use_bias);
    if (W <= 0) return 16;
    size_t bytes = (size_t)b * sizeof(unsigned int) + 16;  // per-sample arrival counters
    if (tn_is_fast_shape(n_layers, dims)) {
        int grid, slots;
        tn_geometry(b, n, grid, slots);
        bytes += (size_t)grid * slots * (size_t)W * sizeof(float);
    }
    return bytes;
}

extern "C" int hp_target_network_backward(int b, int n, int n_layers, const int *dims, int use_bias, const float *weights,
                                          const float *points, long long points_batch_stride, const float *grad_out,
                                          int channels_first, float *grad_weights, float *grad_points, void *workspace,
                                          size_t workspace_bytes, void *stream_v) {
    HP_REQUIRE(b >= 0 && n >= 0, "hp_target_network_backward: negative size (b=%d n=%d)", b, n);
    TNGenArgs g = {};
    TNArgs &a = g.base;
    int rc = tn_fill_offsets(a, n_layers, dims, use_bias, "hp_target_network_backward");
    if (rc != HP_OK) return rc;
    if (b == 0) return HP_OK;
    cudaStream_t stream = (cudaStream_t)stream_v;
    HP_REQUIRE(grad_weights != nullptr, "hp_target_network_backward: grad_weights is null");
    if (n == 0) {
        HP_CUDA(cudaMemsetAsync(grad_weights, 0, (size_t)b * a.W * sizeof(float), stream));
        return HP_OK;
    }
    HP_REQUIRE(weights && points && grad_out, "hp_target_network_backward: null pointer");
    HP_REQUIRE(points_batch_stride == 0 || points_batch_stride >= (long long)n * 3,
               "hp_target_network_backward: points_batch_stride must be 0 (shared cloud) or >= 3n");
    HP_REQUIRE(grad_points == nullptr || points_batch_stride != 0,
               "hp_target_network_backward: grad_points needs per-sample points (points_batch_stride != 0)");
    a.weights = weights, a.points = points, a.pstride = points_batch_stride, a.gout = grad_out;
    a.gweights = grad_weights, a.gpoints = grad_points;
    a.B = b, a.N = n, a.channels_first = channels_first ? 1 : 0;
    if (tn_is_fast_shape(n_layers, dims)) {
        int grid;
        tn_geometry(b, n, grid, a.S);
        const size_t need = hp_target_network_backward_workspace_bytes(b, n, n_layers, dims, use_bias);
        HP_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                   "hp_target_network_backward: workspace null or not 16-byte aligned");
        if (workspace_bytes < need) {
            set_error("hp_target_network_backward: workspace %zu < required %zu bytes", workspace_bytes, need);
            return HP_ERR_WORKSPACE;
        }
        a.counters = reinterpret_cast<unsigned int *>(workspace);
        a.partial = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(workspace) +
                                              (((size_t)b * sizeof(unsigned int) + 15) & ~(size_t)15));
        // the counter region moves with b, so a reused workspace cannot be trusted to be zero there
        HP_CUDA(cudaMemsetAsync(a.counters, 0, (size_t)b * sizeof(unsigned int), stream));
        static SmemAttrCache attr0, attr1, mattr0, mattr1;
        if (g_tn_mode != 1) {
            if (grad_points) {
                HP_CUDA(ensure_dynamic_smem(tn_mma_backward_kernel<true>, TMB_SMEM, mattr1));
                tn_mma_backward_kernel<true><<<(unsigned)grid, TMB_ALL_THREADS, TMB_SMEM, stream>>>(a);
            } else {
                HP_CUDA(ensure_dynamic_smem(tn_mma_backward_kernel<false>, TMB_SMEM, mattr0));
                tn_mma_backward_kernel<false><<<(unsigned)grid, TMB_ALL_THREADS, TMB_SMEM, stream>>>(a);
            }
            HP_LAUNCH_CHECK("tn_mma_backward_kernel");
            return HP_OK;
        }
        if (grad_points) {
            HP_CUDA(ensure_dynamic_smem(tn_backward_kernel<true>, TNB_SMEM, attr1));
            tn_backward_kernel<true><<<(unsigned)grid, TN_THREADS, TNB_SMEM, stream>>>(a);
        } else {
            HP_CUDA(ensure_dynamic_smem(tn_backward_kernel<false>, TNB_SMEM, attr0));
            tn_backward_kernel<false><<<(unsigned)grid, TN_THREADS, TNB_SMEM, stream>>>(a);
        }
        HP_LAUNCH_CHECK("tn_backward_kernel");
        return HP_OK;
    }
    size_t smem_f, smem_b;
    tn_generic_args(g, n_layers, dims, smem_f, smem_b);
    if (smem_b > 200 * 1024) {
        set_error("hp_target_network_backward: widths (sum %d, max %d) need %zu bytes of shared memory (limit 200 KB)", g.sum_dims,
                  g.max_dim, smem_b);
        return HP_ERR_UNSUPPORTED;
    }
    static SmemAttrCache gattr;
    if (smem_b > 48 * 1024) HP_CUDA(ensure_dynamic_smem(tn_generic_backward_kernel, smem_b, gattr));
    tn_generic_backward_kernel<<<b, TNG_THREADS, smem_b, stream>>>(g);
    HP_LAUNCH_CHECK("tn_generic_backward_kernel");
    return HP_OK;
}
