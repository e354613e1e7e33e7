// emd.cu -- approximate earth mover's distance (soft auction) for sm_100a.
//
// Replaces approxmatchkernel / matchcostkernel / matchcostgrad{1,2}kernel of the reference
// (utils/pytorch_structural_losses/approxmatch.cu:34-357) behind hp_approxmatch, hp_matchcost,
// hp_matchcostgrad and adds the match-free hp_emd_cost_pairs used by the metrics path.
//
// The auction runs 9 annealing levels (level = -4^j, j = 7..-1, approxmatch.cu:55-59); each level
// is three all-pairs passes that each need a COMPLETE reduction of the previous one:
//   P1  ratioL[k] = remainL[k] / (1e-9 + sum_l e(k,l) remainR[l])                 (approxmatch.cu:60-93)
//   P2  sumr = remainR[l] * sum_k e(k,l) ratioL[k];  ratioR[l] = min(remainR/(sumr+1e-9),1) remainR[l];
//       remainR[l] = max(0, remainR[l]-sumr)                                       (approxmatch.cu:109-142)
//   P3  w = e(k,l) ratioL[k] ratioR[l];  match[l][k] += w;  remainL[k] = max(0, remainL[k]-sum_l w)
//                                                                                  (approxmatch.cu:161-194)
// with e(k,l) = __expf(level * |x1_k - x2_l|^2).  The reference runs ONE CTA per cloud pair (32 SMs busy
// at B=32).  Here every pass is one launch over ALL cloud pairs: a CTA owns THREADS*RQ rows of one pair
// (and, when a workspace is available, one slice of the columns, so small batches still fill 148 SMs);
// columns are staged SoA in shared memory so one LDS.128 feeds two fp32x2 candidate pairs; the
// distance and the two exponent scalings run as FADD2/FMUL2/FFMA2, the exponential on MUFU.EX2.
// Bound: MUFU (16 ex2/clk/SM) and the FP32 pipe are within ~10% of each other for this mix.
//
// Match-free metrics path.  hp_emd_cost_pairs evaluates every exponential exactly like the reference (same expression in all
// three passes of a level): measured 1.4e-6 worst relative deviation of the cost from the reference extension over 64x64 cloud
// pairs of 2048 points (tests/test_metrics_reference_parity_gpu.py).
// hp_emd_cost_pairs_fast is an opt-in shortcut: P3 of level j and P1 of level j-1 sweep the same (row, column) pairs and P1's sum
// needs nothing of P3 but the row scalar remainL[k], so they run as ONE sweep (emd_fused31_kernel) with ONE ex2 for both -- the
// next level's e' = ex2(d * scale') is evaluated and the current level's e = e'^4 follows by two multiplies (level' = level / 4):
// 3 instead of 4 MUFU operations per pair and level, 19 launches instead of 27, 1.2x faster at B=32, 2048^2 (1.16 vs 1.43 ms)
// when both swept every point.  Since the exact path leaves exhausted points out (emd_compact_kernel) it is the faster one
// (1.18 vs 1.25 ms): the fused sweep keeps every column because its two sums want the lists of two different levels.
// e'^4 deviates from ex2.approx(d * scale) by ~1e-6 relative, P3 then no longer sends exactly what P2 accepted, and the
// nine-level feedback amplifies that to up to 2.1e-5 on the cost (same test) -- above the 1e-5 parity bar, hence not the default.
// (Giving P2 the same e'^4, so that P2 and P3 agree with each other but not with P1, measured WORSE: 1.6e-4.)
//
// Numerics kept from the reference build (verified in its sm_100 SASS):
//   d = fma(dz,dz,fma(dx,dx,dy*dy));  arg = (d * level) * 1.4426950216f;  e = ex2.approx(arg)
//   P1/P2: acc = fma(e, w, acc) in ascending column order;  P3: t = ratioL*e; acc = fma(t, ratioR, acc),
//   match = fma(t, ratioR, match);  IEEE division; level = -powf(4.0f, j) evaluated on the device.
// Deviation: ex2.approx.ftz (results below 2^-126 flush to 0 instead of going denormal): |delta| < 1.2e-38
// per term, below the resolution of every sum it enters.
#include <stdlib.h>

#include "common.cuh"

namespace hp {

constexpr int EMD_CC = 256;  // columns staged per shared-memory chunk

struct EmdArgs {
    const float *first, *second;  // clouds: first [*, n, 3] (xyz1 role), second [*, m, 3] (xyz2 role)
    const int *ia, *ib;           // per-pair cloud index into first / second (nullptr = pair index)
    float *state;                 // [pairs][2(n+m)] = remainL[n] remainR[m] ratioL[n] ratioR[m]
    float *partial;               // [pairs][S][rstride] row partial sums (column-split mode) or nullptr
    float *partial2;              // the same for the second sum of the fused P3 + P1 sweep
    float *match;                 // [pairs][m][n] or nullptr
    float *costpart;              // [pairs][cp_stride] per-CTA cost partials or nullptr
    float *hist;                  // [pairs][9][n+m] per-level ratioL[n] ratioR[m] (match written once at the end) or nullptr
    int n, m, pairs;
    int S, span;                  // column split: slice s covers columns [s*span, min(nc,(s+1)*span))
    int row_tiles, rstride;
    int j;                        // level exponent: level = -4^j
    int level_index;              // 0..8
    int first_level;              // P3: match = w instead of match += w
    int cp_stride;
    int pdl;                      // host side: launch the chain's kernels programmatically dependent (see pdl_trigger)
    // Exhausted points of the second cloud (remainR == 0) drop out of every later pass, see emd_compact_kernel:
    int *active;                  // [pairs][m] indices l with remainR[l] > 0, ascending (nullptr: every point, no compaction)
    int *active_cnt;              // [pairs]
};

__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Programmatic dependent launch along the auction's chain of kernels (init, 27 passes, up to 27 combines, finish): every kernel
// lets its successor become resident as soon as its own CTAs have all started (pdl_trigger, first instruction) and waits for its
// predecessor's completion and memory (pdl_wait) before it touches anything a predecessor writes or still reads -- only the
// point coordinates and the pair indices (inputs of the whole call) are read ahead of the wait.  The successor's launch latency,
// index arithmetic and row-coordinate loads then overlap the predecessor's drain.  Every thread of every kernel executes pdl_wait,
// so no grid of the chain completes before its predecessors; kernels launched normally after the chain wait for all of it.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <int MODE>
__device__ __forceinline__ void emd_row_epilogue(float acc, float *st, int n, int m, int r, float *hist_level = nullptr) {
    float *remainL = st, *remainR = st + n, *ratioL = st + n + m, *ratioR = st + n + m + n;
    if (MODE == 1) {
        const float v = remainL[r] / acc;  // acc already includes the 1e-9 start (approxmatch.cu:68,92)
        ratioL[r] = v;
        if (hist_level) hist_level[r] = v;
    } else if (MODE == 2) {
        const float rem = remainR[r];
        const float sumr = acc * rem;                                   // approxmatch.cu:137
        const float consumption = fminf(rem / (sumr + 1e-9f), 1.0f);    // :138
        ratioR[r] = consumption * rem;                                  // :139
        if (hist_level) hist_level[n + r] = consumption * rem;
        remainR[r] = fmaxf(0.0f, rem - sumr);                           // :140
    } else {
        remainL[r] = fmaxf(0.0f, remainL[r] - acc);                     // :193
    }
}

// MODE 1/3: rows = xyz1 (n), columns = xyz2 (m).  MODE 2: rows = xyz2 (m), columns = xyz1 (n).
template <int MODE, int RQ, int THREADS, bool SPLIT, bool MATCH, bool COST>
__global__ void __launch_bounds__(THREADS) emd_pass_kernel(const EmdArgs a) {
    __shared__ __align__(16) float xs[EMD_CC], ys[EMD_CC], zs[EMD_CC], ws[EMD_CC];
    __shared__ float warp_part[THREADS / 32];
    pdl_trigger();
    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int s = bid % a.S;
    bid /= a.S;
    const int rt = bid % a.row_tiles;
    const int pair = bid / a.row_tiles;
    const int n = a.n, m = a.m;
    const int nr = (MODE == 2) ? m : n, nc = (MODE == 2) ? n : m;
    const size_t c1 = a.ia ? (size_t)a.ia[pair] : (size_t)pair, c2 = a.ib ? (size_t)a.ib[pair] : (size_t)pair;
    const float *__restrict__ X1 = a.first + c1 * n * 3;
    const float *__restrict__ X2 = a.second + c2 * m * 3;
    const float *__restrict__ R = (MODE == 2) ? X2 : X1;
    const float *__restrict__ C = (MODE == 2) ? X1 : X2;
    float *st = a.state + (size_t)pair * 2 * (n + m);
    const float *colw = (MODE == 1) ? st + n : (MODE == 2) ? st + n + m : st + n + m + n;  // remainR | ratioL | ratioR
    // compaction (never with MATCH: that path writes match in place and keeps every column): the points of the second cloud
    // that still have mass, ascending -- the COLUMNS of passes 1 and 3, the ROWS of pass 2
    const bool compact = !MATCH && a.active != nullptr;
    const int *__restrict__ alist = compact ? a.active + (size_t)pair * m : nullptr;
    const float level = -powf(4.0f, (float)a.j);  // approxmatch.cu:56, evaluated on the device like the reference
    // The reference evaluates __expf(level*d) as ex2((d*level)*log2e) with two roundings.  level is a power of two, so
    // d*level is exact and (d*level)*log2e == d*(level*log2e) bit for bit (level*log2e is exact as well): ONE multiply.
    const float scale = level * 1.4426950216293334961f;
    const f32x2 scale2 = pack2(scale, scale);

    f32x2 qx[RQ], qy[RQ], qz[RQ];
    float acc[RQ], rl[RQ], cost[RQ];
    int row[RQ];
    int nc_act = nc;  // columns this pass sweeps: all of them, or the compacted list (passes 1 and 3)
    if (MODE == 2 && compact) {
        // rows = the second cloud's points that still have mass; the list is a result of the chain, so nothing is read ahead
        pdl_wait();
        const int cnt = a.active_cnt[pair];
        if (rt * (THREADS * RQ) >= cnt) return;  // block-uniform: nothing left for this row tile (its rows keep remainR = ratioR = 0)
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
            const int pos = rt * (THREADS * RQ) + q * THREADS + tid;
            row[q] = pos < cnt ? alist[pos] : nr;
        }
    } else {
#pragma unroll
        for (int q = 0; q < RQ; ++q) row[q] = rt * (THREADS * RQ) + q * THREADS + tid;
    }
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
        float x = 0.f, y = 0.f, z = 0.f;
        rl[q] = 0.f;
        if (row[q] < nr) x = __ldg(R + (size_t)row[q] * 3 + 0), y = __ldg(R + (size_t)row[q] * 3 + 1), z = __ldg(R + (size_t)row[q] * 3 + 2);
        qx[q] = pack2(x, x), qy[q] = pack2(y, y), qz[q] = pack2(z, z);
        acc[q] = (MODE == 1 && s == 0) ? 1e-9f : 0.f;  // approxmatch.cu:68 (P1) / :117,169 (P2, P3)
        cost[q] = 0.f;
    }
    if (!(MODE == 2 && compact)) pdl_wait();  // the previous pass (or its combine) is complete: state readable, partial / costpart / match writable
    if (MODE == 3) {
#pragma unroll
        for (int q = 0; q < RQ; ++q)
            if (row[q] < nr) rl[q] = st[n + m + row[q]];  // ratioL[k]
    }
    if (MODE != 2 && compact) nc_act = a.active_cnt[pair];

    // A column with zero weight adds exactly nothing to any sum (fma(e, 0, acc) == acc), so sweeping the compacted list in its
    // ascending order gives the bits of the full sweep.  Column slices (SPLIT) partition the LIST.
    const int c_begin = s * a.span, c_end = min(nc_act, c_begin + a.span);
    for (int c0 = c_begin; c0 < c_end; c0 += EMD_CC) {
        for (int i = tid; i < EMD_CC; i += THREADS) {
            const int p = c0 + i;
            float x = 0.f, y = 0.f, z = 0.f, w = 0.f;  // zero weight: padded columns add exactly nothing
            if (p < c_end) {
                const int c = (MODE != 2 && compact) ? alist[p] : p;
                x = __ldg(C + (size_t)c * 3 + 0), y = __ldg(C + (size_t)c * 3 + 1), z = __ldg(C + (size_t)c * 3 + 2), w = colw[c];
            }
            xs[i] = x, ys[i] = y, zs[i] = z, ws[i] = w;
        }
        __syncthreads();
        const int cnt = min(EMD_CC, c_end - c0);
        const int cnt4 = (cnt + 3) & ~3;
#pragma unroll 2
        for (int c = 0; c < cnt4; c += 4) {
            const ulonglong2 cx = *reinterpret_cast<const ulonglong2 *>(xs + c);
            const ulonglong2 cy = *reinterpret_cast<const ulonglong2 *>(ys + c);
            const ulonglong2 cz = *reinterpret_cast<const ulonglong2 *>(zs + c);
            const float4 w = *reinterpret_cast<const float4 *>(ws + c);
#pragma unroll
            for (int q = 0; q < RQ; ++q) {
                const f32x2 d01 = sqdist_exact2(qx[q], qy[q], qz[q], cx.x, cy.x, cz.x);
                const f32x2 d23 = sqdist_exact2(qx[q], qy[q], qz[q], cx.y, cy.y, cz.y);
                float t0, t1, t2, t3;
                unpack2(mul2(d01, scale2), t0, t1);
                unpack2(mul2(d23, scale2), t2, t3);
                const float e0 = ex2_approx(t0), e1 = ex2_approx(t1), e2 = ex2_approx(t2), e3 = ex2_approx(t3);
                if (MODE != 3) {
                    acc[q] = __fmaf_rn(e0, w.x, acc[q]);
                    acc[q] = __fmaf_rn(e1, w.y, acc[q]);
                    acc[q] = __fmaf_rn(e2, w.z, acc[q]);
                    acc[q] = __fmaf_rn(e3, w.w, acc[q]);
                } else {
                    const float u0 = __fmul_rn(rl[q], e0), u1 = __fmul_rn(rl[q], e1), u2 = __fmul_rn(rl[q], e2), u3 = __fmul_rn(rl[q], e3);
                    acc[q] = __fmaf_rn(u0, w.x, acc[q]);
                    acc[q] = __fmaf_rn(u1, w.y, acc[q]);
                    acc[q] = __fmaf_rn(u2, w.z, acc[q]);
                    acc[q] = __fmaf_rn(u3, w.w, acc[q]);
                    if (MATCH) {
                        if (row[q] < nr) {
                            float *mp = a.match + (size_t)pair * n * m + (size_t)(c0 + c) * n + row[q];
                            const float uu[4] = {u0, u1, u2, u3};
                            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                            for (int v = 0; v < 4; ++v) {
                                if (c + v < cnt) {
                                    const float old = a.first_level ? 0.f : mp[(size_t)v * n];
                                    mp[(size_t)v * n] = __fmaf_rn(uu[v], wv[v], old);
                                }
                            }
                        }
                    }
                    if (COST) {
                        float d0, d1, d2, d3;
                        unpack2(d01, d0, d1);
                        unpack2(d23, d2, d3);
                        cost[q] = __fmaf_rn(__fmul_rn(u0, w.x), sqrt_approx(d0), cost[q]);
                        cost[q] = __fmaf_rn(__fmul_rn(u1, w.y), sqrt_approx(d1), cost[q]);
                        cost[q] = __fmaf_rn(__fmul_rn(u2, w.z), sqrt_approx(d2), cost[q]);
                        cost[q] = __fmaf_rn(__fmul_rn(u3, w.w), sqrt_approx(d3), cost[q]);
                    }
                }
            }
        }
        __syncthreads();
    }

#pragma unroll
    for (int q = 0; q < RQ; ++q) {
        if (row[q] < nr) {
            if (SPLIT) a.partial[((size_t)pair * a.S + s) * a.rstride + row[q]] = acc[q];
            else emd_row_epilogue<MODE>(acc[q], st, n, m, row[q], a.hist ? a.hist + ((size_t)pair * 9 + a.level_index) * (n + m) : nullptr);
        }
    }
    if (COST) {
        float c = 0.f;
#pragma unroll
        for (int q = 0; q < RQ; ++q) c += (row[q] < nr) ? cost[q] : 0.f;
        c = warp_sum(c);
        if ((tid & 31) == 0) warp_part[tid >> 5] = c;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < THREADS / 32; ++i) t += warp_part[i];
            a.costpart[(size_t)pair * a.cp_stride + ((size_t)a.level_index * a.row_tiles + rt) * a.S + s] = t;
        }
    }
}

// Fused sweep of the match-free path: P3 of level `a.j` (weights ratioR, row scalar ratioL, cost) and P1 of level `a.j - 1`
// (weights remainR) over rows = xyz1, columns = xyz2.  Epilogue per row k (non-split): remainL[k] = max(0, remainL[k] - acc3),
// then ratioL[k] = remainL[k] / acc1 with acc1 started at 1e-9 (approxmatch.cu:68,92,193).
template <int RQ, int THREADS, bool SPLIT>
__global__ void __launch_bounds__(THREADS) emd_fused31_kernel(const EmdArgs a) {
    __shared__ __align__(16) float xs[EMD_CC], ys[EMD_CC], zs[EMD_CC], w3s[EMD_CC], w1s[EMD_CC];
    __shared__ float warp_part[THREADS / 32];
    pdl_trigger();
    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int s = bid % a.S;
    bid /= a.S;
    const int rt = bid % a.row_tiles;
    const int pair = bid / a.row_tiles;
    const int n = a.n, m = a.m;
    const size_t c1 = a.ia ? (size_t)a.ia[pair] : (size_t)pair, c2 = a.ib ? (size_t)a.ib[pair] : (size_t)pair;
    const float *__restrict__ R = a.first + c1 * n * 3;
    const float *__restrict__ C = a.second + c2 * m * 3;
    float *st = a.state + (size_t)pair * 2 * (n + m);
    const float *remainR = st + n, *ratioR = st + n + m + n;
    const float level_next = -powf(4.0f, (float)(a.j - 1));  // approxmatch.cu:56, on the device like the reference
    const float scale = level_next * 1.4426950216293334961f;  // exact: the level is a power of two (see emd_pass_kernel)
    const f32x2 scale2 = pack2(scale, scale);

    f32x2 qx[RQ], qy[RQ], qz[RQ];
    float acc3[RQ], acc1[RQ], rl[RQ], cost[RQ];
    int row[RQ];
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
        row[q] = rt * (THREADS * RQ) + q * THREADS + tid;
        float x = 0.f, y = 0.f, z = 0.f;
        rl[q] = 0.f;
        if (row[q] < n) x = __ldg(R + (size_t)row[q] * 3 + 0), y = __ldg(R + (size_t)row[q] * 3 + 1), z = __ldg(R + (size_t)row[q] * 3 + 2);
        qx[q] = pack2(x, x), qy[q] = pack2(y, y), qz[q] = pack2(z, z);
        acc3[q] = 0.f, cost[q] = 0.f;
        acc1[q] = (s == 0) ? 1e-9f : 0.f;
    }
    pdl_wait();
#pragma unroll
    for (int q = 0; q < RQ; ++q)
        if (row[q] < n) rl[q] = st[n + m + row[q]];  // ratioL[k] of the current level
    const int c_begin = s * a.span, c_end = min(m, c_begin + a.span);
    for (int c0 = c_begin; c0 < c_end; c0 += EMD_CC) {
        for (int i = tid; i < EMD_CC; i += THREADS) {
            const int c = c0 + i;
            float x = 0.f, y = 0.f, z = 0.f, w3 = 0.f, w1 = 0.f;  // zero weights: padded columns add exactly nothing
            if (c < c_end) x = __ldg(C + (size_t)c * 3 + 0), y = __ldg(C + (size_t)c * 3 + 1), z = __ldg(C + (size_t)c * 3 + 2), w3 = ratioR[c], w1 = remainR[c];
            xs[i] = x, ys[i] = y, zs[i] = z, w3s[i] = w3, w1s[i] = w1;
        }
        __syncthreads();
        const int cnt4 = (min(EMD_CC, c_end - c0) + 3) & ~3;
#pragma unroll 2
        for (int c = 0; c < cnt4; c += 4) {
            const ulonglong2 cx = *reinterpret_cast<const ulonglong2 *>(xs + c);
            const ulonglong2 cy = *reinterpret_cast<const ulonglong2 *>(ys + c);
            const ulonglong2 cz = *reinterpret_cast<const ulonglong2 *>(zs + c);
            const float4 w3 = *reinterpret_cast<const float4 *>(w3s + c);
            const float4 w1 = *reinterpret_cast<const float4 *>(w1s + c);
            const float w3v[4] = {w3.x, w3.y, w3.z, w3.w}, w1v[4] = {w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int q = 0; q < RQ; ++q) {
                const f32x2 d01 = sqdist_exact2(qx[q], qy[q], qz[q], cx.x, cy.x, cz.x);
                const f32x2 d23 = sqdist_exact2(qx[q], qy[q], qz[q], cx.y, cy.y, cz.y);
                float t[4], d[4];
                unpack2(mul2(d01, scale2), t[0], t[1]);
                unpack2(mul2(d23, scale2), t[2], t[3]);
                unpack2(d01, d[0], d[1]);
                unpack2(d23, d[2], d[3]);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float en = ex2_approx(t[v]);          // next level's e
                    const float e2 = __fmul_rn(en, en);
                    const float ec = __fmul_rn(e2, e2);         // current level's e = en^4
                    acc1[q] = __fmaf_rn(en, w1v[v], acc1[q]);
                    const float u = __fmul_rn(rl[q], ec);
                    acc3[q] = __fmaf_rn(u, w3v[v], acc3[q]);
                    cost[q] = __fmaf_rn(__fmul_rn(u, w3v[v]), sqrt_approx(d[v]), cost[q]);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
        if (row[q] < n) {
            if (SPLIT) {
                a.partial[((size_t)pair * a.S + s) * a.rstride + row[q]] = acc3[q];
                a.partial2[((size_t)pair * a.S + s) * a.rstride + row[q]] = acc1[q];
            } else {
                const float rem = fmaxf(0.0f, st[row[q]] - acc3[q]);  // approxmatch.cu:193
                st[row[q]] = rem;
                st[n + m + row[q]] = rem / acc1[q];                   // :92 (acc1 includes the 1e-9 start, :68)
            }
        }
    }
    float c = 0.f;
#pragma unroll
    for (int q = 0; q < RQ; ++q) c += (row[q] < n) ? cost[q] : 0.f;
    c = warp_sum(c);
    if ((tid & 31) == 0) warp_part[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < THREADS / 32; ++i) t += warp_part[i];
        a.costpart[(size_t)pair * a.cp_stride + ((size_t)a.level_index * a.row_tiles + rt) * a.S + s] = t;
    }
}

// Column-split mode of the fused sweep: fold the S slices of both sums in ascending order, then both row epilogues.
__global__ void emd_combine31_kernel(const EmdArgs a) {
    pdl_trigger();
    pdl_wait();
    const size_t total = (size_t)a.pairs * a.n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int pair = (int)(i / a.n), r = (int)(i % a.n);
        const float *p3 = a.partial + (size_t)pair * a.S * a.rstride + r, *p1 = a.partial2 + (size_t)pair * a.S * a.rstride + r;
        float acc3 = p3[0], acc1 = p1[0];
        for (int s = 1; s < a.S; ++s) acc3 += p3[(size_t)s * a.rstride], acc1 += p1[(size_t)s * a.rstride];
        float *st = a.state + (size_t)pair * 2 * (a.n + a.m);
        const float rem = fmaxf(0.0f, st[r] - acc3);
        st[r] = rem;
        st[a.n + a.m + r] = rem / acc1;
    }
}

// Column-split mode: fold the S slices in ascending order, then the row epilogue.
template <int MODE>
__global__ void emd_combine_kernel(const EmdArgs a) {
    pdl_trigger();
    pdl_wait();
    const int nr = (MODE == 2) ? a.m : a.n;
    const size_t total = (size_t)a.pairs * nr;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int pair = (int)(i / nr), r = (int)(i % nr);
        // compaction: pass 2 left no partial sums for exhausted rows (remainR == 0); their remainR / ratioR stay 0
        if (MODE == 2 && a.active != nullptr && a.state[(size_t)pair * 2 * (a.n + a.m) + a.n + r] <= 0.f) continue;
        const float *p = a.partial + (size_t)pair * a.S * a.rstride + r;
        float acc = p[0];
        for (int s = 1; s < a.S; ++s) acc += p[(size_t)s * a.rstride];
        emd_row_epilogue<MODE>(acc, a.state + (size_t)pair * 2 * (a.n + a.m), a.n, a.m, r,
                               a.hist ? a.hist + ((size_t)pair * 9 + a.level_index) * (a.n + a.m) : nullptr);
    }
}

// Compaction, once per level before pass 1.  A point l of the second cloud whose remainR has reached 0 (the clamp of
// approxmatch.cu:140 hits exactly 0 whenever the point is over-subscribed -- measured on uniform clouds of 2048 points: 10 % of
// the points after the first level, 43 / 65 / 78 / 87 / 93 / 98 / 99 % after the following ones) takes no further part in the
// auction: as a column of pass 1 its weight remainR is 0, as a row of pass 2 its results do not depend on the sum
// (sumr = acc * 0, ratioR = 0, remainR stays 0, :137-140), as a column of pass 3 its weight ratioR is 0.  fma(e, 0, acc) == acc,
// so leaving these points out changes no bit of any sum as long as the others keep their ascending order.  This kernel writes
// the stable list of the points with remainR > 0 per cloud pair and zeroes ratioR (and the recorded per-level ratioR) of the
// others, which pass 2 no longer visits.  One CTA per pair.
constexpr int CP_THREADS = 256;
__global__ void __launch_bounds__(CP_THREADS) emd_compact_kernel(const EmdArgs a) {
    __shared__ int wsum[CP_THREADS / 32];
    pdl_trigger();
    pdl_wait();
    const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, m = a.m;
    float *st = a.state + (size_t)pair * 2 * (n + m);
    const float *remainR = st + n;
    float *ratioR = st + n + m + n;
    float *hist_level = a.hist ? a.hist + ((size_t)pair * 9 + a.level_index) * (n + m) : nullptr;
    int *list = a.active + (size_t)pair * m;
    int base = 0;
    for (int c0 = 0; c0 < m; c0 += CP_THREADS) {  // block-uniform trip count
        const int c = c0 + tid;
        const bool act = c < m && remainR[c] > 0.f;
        if (c < m && !act) {
            ratioR[c] = 0.f;
            if (hist_level) hist_level[n + c] = 0.f;
        }
        const unsigned ball = __ballot_sync(0xffffffffu, act);
        if (lane == 0) wsum[warp] = __popc(ball);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < CP_THREADS / 32; ++w) {
            before += (w < warp) ? wsum[w] : 0;
            total += wsum[w];
        }
        if (act) list[base + before + __popc(ball & ((1u << lane) - 1u))] = c;
        base += total;
        __syncthreads();  // wsum reusable
    }
    if (tid == 0) a.active_cnt[pair] = base;
}

// remainL = multiL, remainR = multiR with the reference's INTEGER division (approxmatch.cu:36-43,49-52)
__global__ void emd_init_kernel(float *state, int pairs, int n, int m) {
    pdl_trigger();
    pdl_wait();
    float multiL, multiR;
    if (n >= m) multiL = 1, multiR = (float)(n / m);
    else multiL = (float)(m / n), multiR = 1;
    const size_t per = (size_t)2 * (n + m), total = (size_t)pairs * (n + m);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t pair = i / (n + m), r = i % (n + m);
        state[pair * per + r] = (r < (size_t)n) ? multiL : multiR;
    }
}

__global__ void emd_cost_finish_kernel(const float *costpart, int cp_stride, int pairs, float *cost) {
    pdl_trigger();
    pdl_wait();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pairs) return;
    float t = 0.f;
    for (int i = 0; i < cp_stride; ++i) t += costpart[(size_t)p * cp_stride + i];  // fixed order
    cost[p] = t;
}

// ---- MatchCost (approxmatch.cu:215-255): HBM-bound read of match; an 8-CTA cluster per cloud pair, the
// per-CTA partials are folded in rank order through distributed shared memory (deterministic, no scratch).
constexpr int MC_CLUSTER = 8;
constexpr int MC_THREADS = 256;

__global__ void __cluster_dims__(MC_CLUSTER, 1, 1) __launch_bounds__(MC_THREADS)
    matchcost_kernel(int b, int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                     const float *__restrict__ match, float *__restrict__ out) {
    __shared__ float warp_part[MC_THREADS / 32];
    __shared__ float cluster_part[MC_CLUSTER];
    unsigned rank;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    // every CTA of the cluster must have started before anyone writes into rank 0's shared memory: arrive now,
    // wait right before the remote store (the main loop runs in between)
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    const int pair = blockIdx.x / MC_CLUSTER;
    const int tid = threadIdx.x;
    const int l_begin = (int)(((long long)m * rank) / MC_CLUSTER), l_end = (int)(((long long)m * (rank + 1)) / MC_CLUSTER);
    const float *X1 = xyz1 + (size_t)pair * n * 3, *X2 = xyz2 + (size_t)pair * m * 3;
    const float *M = match + (size_t)pair * n * m;
    float sum = 0.f;
    for (int k = tid; k < n; k += MC_THREADS) {
        const float x1 = __ldg(X1 + (size_t)k * 3 + 0), y1 = __ldg(X1 + (size_t)k * 3 + 1), z1 = __ldg(X1 + (size_t)k * 3 + 2);
#pragma unroll 8
        for (int l = l_begin; l < l_end; ++l) {
            const float d = sqdist_exact(x1, y1, z1, __ldg(X2 + (size_t)l * 3 + 0), __ldg(X2 + (size_t)l * 3 + 1), __ldg(X2 + (size_t)l * 3 + 2));
            sum = __fmaf_rn(__ldg(M + (size_t)l * n + k), sqrtf(d), sum);  // approxmatch.cu:238-239 (IEEE sqrtf)
        }
    }
    sum = warp_sum(sum);
    if ((tid & 31) == 0) warp_part[tid >> 5] = sum;
    __syncthreads();
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");  // all CTAs of the cluster are running
    if (tid == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < MC_THREADS / 32; ++i) t += warp_part[i];
        // write my partial into rank 0's cluster_part[rank]
        uint32_t local = smem_u32(&cluster_part[rank]), remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0));
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(t) : "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (rank == 0 && tid == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < MC_CLUSTER; ++i) t += cluster_part[i];
        out[pair] = t;
    }
}

// ---- MatchCostGrad (approxmatch.cu:260-322) ----------------------------------------------------------
// grad1[k] = sum_l match[l][k] (x1_k - x2_l) rsqrt(max(d,1e-20)): thread per k, ascending l (the reference's order)
__global__ void matchcostgrad1_kernel(int b, int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                                      const float *__restrict__ match, float *__restrict__ grad1) {
    extern __shared__ __align__(16) float sh[];  // xyz2 tile [LT*3]
    constexpr int LT = 512;
    const int pair = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const float *X2 = xyz2 + (size_t)pair * m * 3;
    const float *M = match + (size_t)pair * n * m;
    float x1 = 0.f, y1 = 0.f, z1 = 0.f;
    if (k < n) {
        const float *p = xyz1 + ((size_t)pair * n + k) * 3;
        x1 = __ldg(p), y1 = __ldg(p + 1), z1 = __ldg(p + 2);
    }
    float gx = 0.f, gy = 0.f, gz = 0.f;
    for (int l0 = 0; l0 < m; l0 += LT) {
        const int cnt = min(LT, m - l0);
        for (int i = threadIdx.x; i < cnt * 3; i += blockDim.x) sh[i] = __ldg(X2 + (size_t)l0 * 3 + i);
        __syncthreads();
        if (k < n) {
#pragma unroll 8
            for (int l = 0; l < cnt; ++l) {
                const float ex = x1 - sh[l * 3 + 0], ey = y1 - sh[l * 3 + 1], ez = z1 - sh[l * 3 + 2];
                const float s2 = __fmaf_rn(ez, ez, __fmaf_rn(ex, ex, __fmul_rn(ey, ey)));
                const float d = __ldg(M + (size_t)(l0 + l) * n + k) * rsqrtf(fmaxf(s2, 1e-20f));
                gx = __fmaf_rn(ex, d, gx), gy = __fmaf_rn(ey, d, gy), gz = __fmaf_rn(ez, d, gz);
            }
        }
        __syncthreads();
    }
    if (k < n) {
        float *g = grad1 + ((size_t)pair * n + k) * 3;
        g[0] = gx, g[1] = gy, g[2] = gz;
    }
}

// grad2[l] = sum_k match[l][k] (x2_l - x1_k) rsqrt(max(d,1e-20)): one warp per l, lanes stride k, fixed shuffle tree
__global__ void matchcostgrad2_kernel(int b, int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                                      const float *__restrict__ match, float *__restrict__ grad2) {
    const int pair = blockIdx.y;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= m) return;
    const int l = warp;
    const float *X1 = xyz1 + (size_t)pair * n * 3;
    const float *p2 = xyz2 + ((size_t)pair * m + l) * 3;
    const float x2 = __ldg(p2), y2 = __ldg(p2 + 1), z2 = __ldg(p2 + 2);
    const float *M = match + (size_t)pair * n * m + (size_t)l * n;
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll 4
    for (int k = lane; k < n; k += 32) {
        const float ex = x2 - __ldg(X1 + (size_t)k * 3 + 0), ey = y2 - __ldg(X1 + (size_t)k * 3 + 1), ez = z2 - __ldg(X1 + (size_t)k * 3 + 2);
        const float s2 = __fmaf_rn(ez, ez, __fmaf_rn(ex, ex, __fmul_rn(ey, ey)));
        const float d = __ldg(M + k) * rsqrtf(fmaxf(s2, 1e-20f));
        gx = __fmaf_rn(ex, d, gx), gy = __fmaf_rn(ey, d, gy), gz = __fmaf_rn(ez, d, gz);
    }
    gx = warp_sum(gx), gy = warp_sum(gy), gz = warp_sum(gz);
    if (lane == 0) {
        float *g = grad2 + ((size_t)pair * m + l) * 3;
        g[0] = gx, g[1] = gy, g[2] = gz;
    }
}


// ---- match written ONCE (hp_approxmatch_ws): match[l][k] = sum over the 9 levels, in level order, of
//      fma(ratioL_lvl[k] * e_lvl(k,l), ratioR_lvl[l], .)  -- the same fma chain the reference builds with nine
//      read-modify-write sweeps over the [b,m,n] matrix (approxmatch.cu:181-188), from the per-level ratios the
//      auction recorded.  9 more ex2 per pair, but 4*n*m bytes of HBM traffic per cloud instead of 9*2*4*n*m.
constexpr int MW_THREADS = 128, MW_RQ = 2, MW_LC = 32;  // k rows per CTA = 256, l columns staged per chunk = 32

__global__ void __launch_bounds__(MW_THREADS) emd_match_write_kernel(const EmdArgs a, int l_span) {
    __shared__ float cx[MW_LC], cy[MW_LC], cz[MW_LC], rr[9][MW_LC];
    const int tid = threadIdx.x;
    const int n = a.n, m = a.m;
    const int ktiles = (n + MW_THREADS * MW_RQ - 1) / (MW_THREADS * MW_RQ);
    int bid = blockIdx.x;
    const int kt = bid % ktiles;
    bid /= ktiles;
    const int nls = (m + l_span - 1) / l_span;
    const int ls = bid % nls;
    const int pair = bid / nls;
    const float *__restrict__ X1 = a.first + (size_t)pair * n * 3;
    const float *__restrict__ X2 = a.second + (size_t)pair * m * 3;
    const float *__restrict__ H = a.hist + (size_t)pair * 9 * (n + m);
    float *__restrict__ M = a.match + (size_t)pair * n * m;
    float scale[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) scale[i] = -powf(4.0f, (float)(7 - i)) * 1.4426950216293334961f;
    float qx[MW_RQ], qy[MW_RQ], qz[MW_RQ], rl[MW_RQ][9];
    int k[MW_RQ];
#pragma unroll
    for (int q = 0; q < MW_RQ; ++q) {
        k[q] = kt * (MW_THREADS * MW_RQ) + q * MW_THREADS + tid;
        const bool ok = k[q] < n;
        qx[q] = ok ? __ldg(X1 + (size_t)k[q] * 3 + 0) : 0.f, qy[q] = ok ? __ldg(X1 + (size_t)k[q] * 3 + 1) : 0.f;
        qz[q] = ok ? __ldg(X1 + (size_t)k[q] * 3 + 2) : 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) rl[q][i] = ok ? __ldg(H + (size_t)i * (n + m) + k[q]) : 0.f;
    }
    const int l_begin = ls * l_span, l_end = min(m, l_begin + l_span);
    for (int l0 = l_begin; l0 < l_end; l0 += MW_LC) {
        __syncthreads();
        for (int i = tid; i < MW_LC * 12; i += MW_THREADS) {
            const int c = i % MW_LC, f = i / MW_LC, l = l0 + c;
            float v = 0.f;
            if (l < l_end) v = (f < 3) ? __ldg(X2 + (size_t)l * 3 + f) : __ldg(H + (size_t)(f - 3) * (n + m) + n + l);
            if (f == 0) cx[c] = v;
            else if (f == 1) cy[c] = v;
            else if (f == 2) cz[c] = v;
            else rr[f - 3][c] = v;
        }
        __syncthreads();
        const int cnt = min(MW_LC, l_end - l0);
        for (int c = 0; c < cnt; ++c) {
#pragma unroll
            for (int q = 0; q < MW_RQ; ++q) {
                const float d = sqdist_exact(qx[q], qy[q], qz[q], cx[c], cy[c], cz[c]);
                float mv = 0.f;
#pragma unroll
                for (int i = 0; i < 9; ++i) mv = __fmaf_rn(__fmul_rn(rl[q][i], ex2_approx(__fmul_rn(d, scale[i]))), rr[i][c], mv);
                if (k[q] < n) M[(size_t)(l0 + c) * n + k[q]] = mv;
            }
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------
// A launch of the auction's chain, programmatically dependent when `pdl` (see pdl_trigger / pdl_wait).  Only auctions whose
// passes fill the GPU take it (chain_pdl): early residency packs a SMALL grid onto the few SMs that happen to have room while its
// predecessor runs, and successors of successors pile up behind it -- measured on the un-split ApproxMatch at B=32 (512 CTAs of 64
// threads per pass): 2.24 -> 3.7 ms with the attribute, against 1.52 -> 1.43 ms for the column-split cost path (1024 CTAs of 128).
template <typename... KArgs, typename... Args>
static cudaError_t launch_chain(bool pdl, void (*kern)(KArgs...), unsigned grid, unsigned block, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(block), cfg.dynamicSmemBytes = 0, cfg.stream = stream;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs, cfg.numAttrs = pdl ? 1 : 0;
#ifdef HP_BENCH_BUILD
    {
        static int off = -1;  // A/B: plain stream-ordered launches
        if (off < 0) {
            const char *e = getenv("HP_EMD_NO_PDL");
            off = (e && atoi(e)) ? 1 : 0;
        }
        if (off) cfg.numAttrs = 0;
    }
#endif
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// true when launch_pass_auto will pick a 128-thread tile for every pass with at least four CTAs per SM
static bool chain_pdl(const EmdArgs &a) {
    const int small_ = a.n < a.m ? a.n : a.m;
    return (long long)a.pairs * ((small_ + 511) / 512) * a.S >= (long long)sm_count() * 4;
}

template <int MODE, int RQ, int THREADS, bool SPLIT, bool MATCH, bool COST>
static int launch_pass(EmdArgs a, cudaStream_t stream) {
    const int nr = (MODE == 2) ? a.m : a.n;
    a.row_tiles = (nr + THREADS * RQ - 1) / (THREADS * RQ);
    const long long grid = (long long)a.pairs * a.row_tiles * a.S;
    HP_REQUIRE(grid <= 0x7fffffffLL, "emd: grid too large (%lld CTAs); split the batch", grid);
    HP_CUDA(launch_chain(a.pdl != 0, emd_pass_kernel<MODE, RQ, THREADS, SPLIT, MATCH, COST>, (unsigned)grid, THREADS, stream, a));
    return HP_OK;
}

// Tile shape by available row parallelism: big tiles (RQ=8: shared-memory traffic per ex2 is lowest) when
// the batch alone fills the GPU, smaller ones otherwise.
template <int MODE, bool SPLIT, bool MATCH, bool COST>
static int launch_pass_auto(const EmdArgs &a, cudaStream_t stream) {
    const int nr = (MODE == 2) ? a.m : a.n;
    const long long want = (long long)sm_count() * 4;  // CTAs
    auto ctas = [&](int tile) { return (long long)a.pairs * ((nr + tile - 1) / tile) * a.S; };
    if (ctas(128 * 8) >= want) return launch_pass<MODE, 8, 128, SPLIT, MATCH, COST>(a, stream);
    if (ctas(128 * 4) >= want) return launch_pass<MODE, 4, 128, SPLIT, MATCH, COST>(a, stream);
    if (ctas(64 * 4) >= want) return launch_pass<MODE, 4, 64, SPLIT, MATCH, COST>(a, stream);
    return launch_pass<MODE, 2, 64, SPLIT, MATCH, COST>(a, stream);
}

static int row_tiles_auto(int pairs, int nr, int S) {
    const long long want = (long long)sm_count() * 4;
    auto ctas = [&](int tile) { return (long long)pairs * ((nr + tile - 1) / tile) * S; };
    int tile = 64 * 2;
    if (ctas(128 * 8) >= want) tile = 128 * 8;
    else if (ctas(128 * 4) >= want) tile = 128 * 4;
    else if (ctas(64 * 4) >= want) tile = 64 * 4;
    return (nr + tile - 1) / tile;
}

template <bool SPLIT, bool MATCH, bool COST>
static int run_auction(EmdArgs a, cudaStream_t stream) {
    const int blocks = sm_count() * 4;
    a.pdl = chain_pdl(a) ? 1 : 0;
    HP_CUDA(launch_chain(a.pdl != 0, emd_init_kernel, blocks, 256, stream, a.state, a.pairs, a.n, a.m));
    const int span_rows_n = a.S > 1 ? ((a.n + a.S - 1) / a.S + EMD_CC - 1) / EMD_CC * EMD_CC : a.n;  // columns = xyz1 (P2)
    const int span_rows_m = a.S > 1 ? ((a.m + a.S - 1) / a.S + EMD_CC - 1) / EMD_CC * EMD_CC : a.m;  // columns = xyz2 (P1, P3)
    int li = 0;
    for (int j = 7; j > -2; --j, ++li) {  // approxmatch.cu:55
        a.j = j;
        a.level_index = li;
        a.first_level = (li == 0);
        int rc;
        if (!MATCH && a.active != nullptr) HP_CUDA(launch_chain(a.pdl != 0, emd_compact_kernel, (unsigned)a.pairs, CP_THREADS, stream, a));
        a.span = span_rows_m;
        if ((rc = launch_pass_auto<1, SPLIT, false, false>(a, stream)) != HP_OK) return rc;
        if (SPLIT) {
            HP_CUDA(launch_chain(a.pdl != 0, emd_combine_kernel<1>, blocks, 256, stream, a));
            HP_LAUNCH_CHECK("emd_combine_kernel<1>");
        }
        a.span = span_rows_n;
        if ((rc = launch_pass_auto<2, SPLIT, false, false>(a, stream)) != HP_OK) return rc;
        if (SPLIT) {
            HP_CUDA(launch_chain(a.pdl != 0, emd_combine_kernel<2>, blocks, 256, stream, a));
            HP_LAUNCH_CHECK("emd_combine_kernel<2>");
        }
        a.span = span_rows_m;
        if ((rc = launch_pass_auto<3, SPLIT, MATCH, COST>(a, stream)) != HP_OK) return rc;
        if (SPLIT) {
            HP_CUDA(launch_chain(a.pdl != 0, emd_combine_kernel<3>, blocks, 256, stream, a));
            HP_LAUNCH_CHECK("emd_combine_kernel<3>");
        }
    }
    return HP_OK;
}

template <int RQ, int THREADS, bool SPLIT>
static int launch_fused31(EmdArgs a, cudaStream_t stream) {
    a.row_tiles = (a.n + THREADS * RQ - 1) / (THREADS * RQ);
    const long long grid = (long long)a.pairs * a.row_tiles * a.S;
    HP_REQUIRE(grid <= 0x7fffffffLL, "emd: grid too large (%lld CTAs); split the batch", grid);
    HP_CUDA(launch_chain(a.pdl != 0, emd_fused31_kernel<RQ, THREADS, SPLIT>, (unsigned)grid, THREADS, stream, a));
    HP_LAUNCH_CHECK("emd_fused31_kernel");
    return HP_OK;
}
// same tile rule as launch_pass_auto<3> / row_tiles_auto (the cost partial layout depends on it)
template <bool SPLIT>
static int launch_fused31_auto(const EmdArgs &a, cudaStream_t stream) {
    const long long want = (long long)sm_count() * 4;
    auto ctas = [&](int tile) { return (long long)a.pairs * ((a.n + tile - 1) / tile) * a.S; };
    if (ctas(128 * 8) >= want) return launch_fused31<8, 128, SPLIT>(a, stream);
    if (ctas(128 * 4) >= want) return launch_fused31<4, 128, SPLIT>(a, stream);
    if (ctas(64 * 4) >= want) return launch_fused31<4, 64, SPLIT>(a, stream);
    return launch_fused31<2, 64, SPLIT>(a, stream);
}

// The match-free auction with P3(level) + P1(next level) fused: P1(first level); then per level P2, fused sweep (the last
// level: plain P3 with the cost).  19 launches (+ one combine per launch in column-split mode).
template <bool SPLIT>
static int run_auction_fused(EmdArgs a, cudaStream_t stream) {
    const int blocks = sm_count() * 4;
    a.pdl = chain_pdl(a) ? 1 : 0;
    HP_CUDA(launch_chain(a.pdl != 0, emd_init_kernel, blocks, 256, stream, a.state, a.pairs, a.n, a.m));
    const int span_rows_n = a.S > 1 ? ((a.n + a.S - 1) / a.S + EMD_CC - 1) / EMD_CC * EMD_CC : a.n;  // columns = xyz1 (P2)
    const int span_rows_m = a.S > 1 ? ((a.m + a.S - 1) / a.S + EMD_CC - 1) / EMD_CC * EMD_CC : a.m;  // columns = xyz2 (P1, P3)
    int rc, li = 0;
    a.j = 7, a.level_index = 0, a.first_level = 1, a.span = span_rows_m;
    if ((rc = launch_pass_auto<1, SPLIT, false, false>(a, stream)) != HP_OK) return rc;
    if (SPLIT) {
        HP_CUDA(launch_chain(a.pdl != 0, emd_combine_kernel<1>, blocks, 256, stream, a));
        HP_LAUNCH_CHECK("emd_combine_kernel<1>");
    }
    for (int j = 7; j > -2; --j, ++li) {  // approxmatch.cu:55
        a.j = j, a.level_index = li, a.first_level = (li == 0);
        a.span = span_rows_n;
        if ((rc = launch_pass_auto<2, SPLIT, false, false>(a, stream)) != HP_OK) return rc;
        if (SPLIT) {
            HP_CUDA(launch_chain(a.pdl != 0, emd_combine_kernel<2>, blocks, 256, stream, a));
            HP_LAUNCH_CHECK("emd_combine_kernel<2>");
        }
        a.span = span_rows_m;
        if (j > -1) {
            if ((rc = launch_fused31_auto<SPLIT>(a, stream)) != HP_OK) return rc;
            if (SPLIT) {
                HP_CUDA(launch_chain(a.pdl != 0, emd_combine31_kernel, blocks, 256, stream, a));
                HP_LAUNCH_CHECK("emd_combine31_kernel");
            }
        } else {  // last level: nothing follows, plain P3 with the cost (remainL is not needed any more)
            if ((rc = launch_pass_auto<3, SPLIT, false, true>(a, stream)) != HP_OK) return rc;
        }
    }
    return HP_OK;
}

// Column split for the workspace-backed path.  Work items = pairs x row tiles x column slices, all of equal cost; the
// slice count is chosen so that the items spread evenly over the SMs (items / (ceil(items/SMs)*SMs) close to 1: the
// reference shape B=32, 2048^2 gets 8 slices -> 1024 items = 6.9 per SM instead of 512 = 3.5 per SM), preferring fewer
// slices on ties.  Returns the number of NON-EMPTY slices for spans rounded up to whole shared-memory chunks.
static int choose_split(int pairs, int n, int m) {
    const int sms = sm_count();
    const int big = n > m ? n : m, small_ = n > m ? m : n;
    const int maxS = (big + EMD_CC - 1) / EMD_CC;
    const long long row_tiles = (small_ + 511) / 512;  // the 128x4 row tile launch_pass_auto prefers
    double best_eff = -1.0;
    int best = 1;
    for (int S = 1; S <= maxS; ++S) {
        const int span = ((big + S - 1) / S + EMD_CC - 1) / EMD_CC * EMD_CC;
        const int real = (big + span - 1) / span;
        if (real != S) continue;  // rounding left an empty slice: skip this count
        const long long items = (long long)pairs * row_tiles * S;
        const long long waves = (items + sms - 1) / sms;
        double eff = (double)items / (double)(waves * sms);
        if (items < 2LL * sms) eff *= 0.5 * (double)items / (2.0 * sms) + 0.5;  // too few items: poor latency hiding
        if (eff > best_eff + 0.02) best_eff = eff, best = S;
        if (items >= 16LL * sms) break;
    }
    return best;
}

}  // namespace hp

using namespace hp;

extern "C" int hp_approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, float *temp,
                              void *stream) {
    HP_REQUIRE(b >= 0 && n >= 0 && m >= 0, "hp_approxmatch: negative size (b=%d n=%d m=%d)", b, n, m);
    if (b == 0 || (n == 0 && m == 0)) return HP_OK;
    HP_REQUIRE(n > 0 && m > 0, "hp_approxmatch: one point set is empty (n=%d m=%d): the reference divides n/m", n, m);
    HP_REQUIRE(xyz1 && xyz2 && match && temp, "hp_approxmatch: null pointer");
    EmdArgs a = {};
    a.first = xyz1, a.second = xyz2, a.ia = nullptr, a.ib = nullptr;
    a.state = temp, a.partial = nullptr, a.match = match, a.costpart = nullptr;
    a.n = n, a.m = m, a.pairs = b, a.S = 1, a.rstride = 0, a.cp_stride = 0;
    return run_auction<false, true, false>(a, (cudaStream_t)stream);
}

extern "C" size_t hp_approxmatch_workspace_bytes(int b, int n, int m) {
    if (b <= 0 || n <= 0 || m <= 0) return 16;
    // per-level ratioL | ratioR, then the compaction list of every batch element and its length
    return ((size_t)b * 9 * ((size_t)n + m) + (size_t)b * ((size_t)m + 1)) * sizeof(float) + 64;
}

extern "C" int hp_approxmatch_ws(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, float *temp,
                                 void *workspace, size_t workspace_bytes, void *stream_v) {
    HP_REQUIRE(b >= 0 && n >= 0 && m >= 0, "hp_approxmatch_ws: negative size (b=%d n=%d m=%d)", b, n, m);
    if (b == 0 || (n == 0 && m == 0)) return HP_OK;
    HP_REQUIRE(n > 0 && m > 0, "hp_approxmatch_ws: one point set is empty (n=%d m=%d): the reference divides n/m", n, m);
    HP_REQUIRE(xyz1 && xyz2 && match && temp && workspace, "hp_approxmatch_ws: null pointer");
    HP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "hp_approxmatch_ws: workspace must be 16-byte aligned");
    if (workspace_bytes < hp_approxmatch_workspace_bytes(b, n, m)) {
        set_error("hp_approxmatch_ws: workspace %zu < required %zu bytes", workspace_bytes, hp_approxmatch_workspace_bytes(b, n, m));
        return HP_ERR_WORKSPACE;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    EmdArgs a = {};
    a.first = xyz1, a.second = xyz2, a.ia = nullptr, a.ib = nullptr;
    a.state = temp;  // remainL | remainR | ratioL | ratioR, returned like the reference's temp
    // No column split here: every row sum runs over ALL columns in ascending order like the reference's thread does.
    // The 9-level feedback amplifies a different association to ~1e-3 on the gradients (measured), which would miss
    // the 1e-5 bar; the match-free metrics path (hp_emd_cost_pairs) may split because the COST stays within 1e-6.
    a.partial = nullptr;
    a.hist = reinterpret_cast<float *>(workspace);
    a.active = reinterpret_cast<int *>(a.hist + (size_t)b * 9 * ((size_t)n + m));
    a.active_cnt = a.active + (size_t)b * m;
    a.match = match, a.costpart = nullptr;
    a.n = n, a.m = m, a.pairs = b, a.S = 1, a.rstride = 0, a.cp_stride = 0;
    int rc = run_auction<false, false, false>(a, stream);
    if (rc != HP_OK) return rc;
    // one sweep over match: k tiles x l slices sized for ~8 CTAs per SM
    const int ktiles = (n + MW_THREADS * MW_RQ - 1) / (MW_THREADS * MW_RQ);
    long long want = (long long)sm_count() * 8;
    long long slices = (want + (long long)b * ktiles - 1) / ((long long)b * ktiles);
    const long long max_slices = (m + MW_LC - 1) / MW_LC;
    if (slices > max_slices) slices = max_slices;
    if (slices < 1) slices = 1;
    int l_span = (int)(((m + slices - 1) / slices + MW_LC - 1) / MW_LC * MW_LC);
    const long long nls = (m + l_span - 1) / l_span;
    const long long grid = (long long)b * nls * ktiles;
    HP_REQUIRE(grid <= 0x7fffffffLL, "hp_approxmatch_ws: grid too large; split the batch");
    emd_match_write_kernel<<<(unsigned)grid, MW_THREADS, 0, stream>>>(a, l_span);
    HP_LAUNCH_CHECK("emd_match_write_kernel");
    return HP_OK;
}

extern "C" size_t hp_emd_cost_workspace_bytes(int pairs, int n, int m) {
    if (pairs <= 0 || n <= 0 || m <= 0) return 16;
    const int S = choose_split(pairs, n, m);
    const size_t big = (size_t)(n > m ? n : m);
    const size_t state = (size_t)pairs * 2 * ((size_t)n + m);
    const size_t partial = S > 1 ? (size_t)2 * pairs * S * big : 0;  // two sums in the fused P3 + P1 sweep
    const int rt_max = (n + 127) / 128;  // upper bound: the smallest row tile is 64 threads x 2 rows
    const size_t costpart = (size_t)pairs * 9 * rt_max * S;
    const size_t active = (size_t)pairs * ((size_t)m + 1);  // compaction list and its length per pair (ints)
    return (state + partial + costpart + active) * sizeof(float) + 64;
}

static int emd_cost_pairs_impl(int pairs, int n, int m, const float *first, const int *ia, const float *second, const int *ib,
                               float *cost, void *workspace, size_t workspace_bytes, void *stream_v, bool fused) {
    HP_REQUIRE(pairs >= 0 && n >= 0 && m >= 0, "hp_emd_cost_pairs: negative size (pairs=%d n=%d m=%d)", pairs, n, m);
    if (pairs == 0) return HP_OK;
    HP_REQUIRE(n > 0 && m > 0, "hp_emd_cost_pairs: empty point set (n=%d m=%d)", n, m);
    HP_REQUIRE(first && second && cost && workspace, "hp_emd_cost_pairs: null pointer");
    HP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "hp_emd_cost_pairs: workspace must be 16-byte aligned");
    if (workspace_bytes < hp_emd_cost_workspace_bytes(pairs, n, m)) {
        set_error("hp_emd_cost_pairs: workspace %zu < required %zu bytes", workspace_bytes, hp_emd_cost_workspace_bytes(pairs, n, m));
        return HP_ERR_WORKSPACE;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    const int S = choose_split(pairs, n, m);
    const size_t big = (size_t)(n > m ? n : m);
    float *ws = reinterpret_cast<float *>(workspace);
    EmdArgs a = {};
    a.first = first, a.second = second, a.ia = ia, a.ib = ib;
    a.state = ws;
    ws += (size_t)pairs * 2 * ((size_t)n + m);
    a.partial = S > 1 ? ws : nullptr;
    ws += S > 1 ? (size_t)pairs * S * big : 0;
    a.partial2 = S > 1 ? ws : nullptr;
    ws += S > 1 ? (size_t)pairs * S * big : 0;
    a.costpart = ws;
    ws += (size_t)pairs * 9 * ((n + 127) / 128) * S;  // the bound hp_emd_cost_workspace_bytes reserves
    if (!fused) {  // the opt-in fused sweep keeps every column (its two sums want different lists)
        a.active = reinterpret_cast<int *>(ws);
        a.active_cnt = a.active + (size_t)pairs * m;
    }
    a.match = nullptr;
    a.n = n, a.m = m, a.pairs = pairs, a.S = S, a.rstride = (int)big;
    const int rt3 = row_tiles_auto(pairs, n, S);  // the row tiling P3 will use (same rule as launch_pass_auto<3>)
    a.cp_stride = 9 * rt3 * S;
    int rc;
    if (fused) rc = (S > 1) ? run_auction_fused<true>(a, stream) : run_auction_fused<false>(a, stream);
    else rc = (S > 1) ? run_auction<true, false, true>(a, stream) : run_auction<false, false, true>(a, stream);
    if (rc != HP_OK) return rc;
    HP_CUDA(launch_chain(chain_pdl(a), emd_cost_finish_kernel, (unsigned)((pairs + 127) / 128), 128, stream, (const float *)a.costpart, a.cp_stride, pairs, cost));
    HP_LAUNCH_CHECK("emd_cost_finish_kernel");
    return HP_OK;
}

extern "C" int hp_emd_cost_pairs(int pairs, int n, int m, const float *first, const int *ia, const float *second,
                                 const int *ib, float *cost, void *workspace, size_t workspace_bytes, void *stream) {
    return emd_cost_pairs_impl(pairs, n, m, first, ia, second, ib, cost, workspace, workspace_bytes, stream, false);
}

extern "C" int hp_emd_cost_pairs_fast(int pairs, int n, int m, const float *first, const int *ia, const float *second,
                                      const int *ib, float *cost, void *workspace, size_t workspace_bytes, void *stream) {
    return emd_cost_pairs_impl(pairs, n, m, first, ia, second, ib, cost, workspace, workspace_bytes, stream, true);
}

extern "C" int hp_matchcost(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match, float *out,
                            void *stream) {
    HP_REQUIRE(b >= 0 && n >= 0 && m >= 0, "hp_matchcost: negative size (b=%d n=%d m=%d)", b, n, m);
    if (b == 0) return HP_OK;
    HP_REQUIRE(out != nullptr, "hp_matchcost: null output");
    if (n == 0 || m == 0) {
        HP_CUDA(cudaMemsetAsync(out, 0, (size_t)b * sizeof(float), (cudaStream_t)stream));
        return HP_OK;
    }
    HP_REQUIRE(xyz1 && xyz2 && match, "hp_matchcost: null pointer");
    matchcost_kernel<<<b * MC_CLUSTER, MC_THREADS, 0, (cudaStream_t)stream>>>(b, n, m, xyz1, xyz2, match, out);
    HP_LAUNCH_CHECK("matchcost_kernel");
    return HP_OK;
}

extern "C" int hp_matchcostgrad(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match,
                                float *grad1, float *grad2, void *stream_v) {
    HP_REQUIRE(b >= 0 && n >= 0 && m >= 0, "hp_matchcostgrad: negative size (b=%d n=%d m=%d)", b, n, m);
    if (b == 0 || (n == 0 && m == 0)) return HP_OK;
    HP_REQUIRE(b <= 65535, "hp_matchcostgrad: batch %d > 65535; split the batch", b);
    cudaStream_t stream = (cudaStream_t)stream_v;
    HP_REQUIRE(grad1 && grad2, "hp_matchcostgrad: null output");
    if (n == 0 || m == 0) {
        if (n) HP_CUDA(cudaMemsetAsync(grad1, 0, (size_t)b * n * 3 * sizeof(float), stream));
        if (m) HP_CUDA(cudaMemsetAsync(grad2, 0, (size_t)b * m * 3 * sizeof(float), stream));
        return HP_OK;
    }
    HP_REQUIRE(xyz1 && xyz2 && match, "hp_matchcostgrad: null pointer");
    matchcostgrad1_kernel<<<dim3((n + 127) / 128, b), 128, 512 * 3 * sizeof(float), stream>>>(b, n, m, xyz1, xyz2, match, grad1);
    HP_LAUNCH_CHECK("matchcostgrad1_kernel");
    matchcostgrad2_kernel<<<dim3((m * 32 + 255) / 256, b), 256, 0, stream>>>(b, n, m, xyz1, xyz2, match, grad2);
    HP_LAUNCH_CHECK("matchcostgrad2_kernel");
    return HP_OK;
}
