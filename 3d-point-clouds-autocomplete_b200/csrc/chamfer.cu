// chamfer.cu -- nearest-neighbour (Chamfer) forward / backward for sm_100a.
//
// Replaces NmDistanceKernel / NmDistanceGradKernel of the reference
// (utils/pytorch_structural_losses/nndistance.cu:8-160) behind hp_nndistance,
// hp_nndistancegrad, hp_chamfer_forward, hp_chamfer_backward (include/hp_b200.h).
//
// Forward design (one launch, both directions):
//   * a CTA owns THREADS*RQ query points of one cloud and one direction; every thread keeps
//     RQ queries in registers (each coordinate duplicated into an fp32x2 pair);
//   * the candidate cloud is brought into shared memory with ONE 1-D bulk-TMA copy per
//     chunk (cp.async.bulk, mbarrier completion), then re-laid out SoA so that an LDS.128
//     delivers the same coordinate of 4 consecutive candidates = two ready-made fp32x2 pairs;
//   * the K=3 contraction stays on the FP32 pipe: FADD2/FMUL2/FFMA2 evaluate two candidates
//     per instruction with the reference's exact association, FMNMX3 folds two distances
//     into the running minimum per instruction;
//   * only the minimum VALUE is tracked in the hot loop; the sub-chunk (16 candidates) in
//     which the running minimum last strictly improved is remembered, and the argmin is
//     recovered afterwards by re-evaluating that one sub-chunk and taking the lowest index
//     with d == min (bit-exact same arithmetic) -> identical to the reference's
//     "strict <, lowest index wins" rule (nndistance.cu:32-64,117-125);
//   * optional fused loss: per-CTA partial sums in fixed order, last CTA folds them.
#include <stdlib.h>

#include "common.cuh"

namespace hp {

constexpr int NN_CH = 16;  // candidates per index-tracking sub-chunk

struct NNArgs {
    const float *set[2];  // set[0] = xyz1 [b,n,3], set[1] = xyz2 [b,m,3]
    float *dist[2];
    int *idx[2];
    int npts[2];
    int tiles[2];  // query tiles per cloud, per direction
    int b;
    float *loss;            // nullptr -> no fused loss
    float *partial;         // [gridDim.x]
    unsigned int *counter;  // zero on entry, zero on exit
};

template <int THREADS, int RQ, int MC>
__global__ void __launch_bounds__(THREADS) nn_fwd_kernel(const NNArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // One [MC*3] buffer: the bulk copy lands the chunk AoS (xyzxyz...), then it is transposed IN PLACE
    // (through registers) to SoA xs|ys|zs so that an LDS.128 yields two ready-made fp32x2 candidate pairs.
    float *stage = reinterpret_cast<float *>(smem_raw);
    float *xs = stage;
    float *ys = xs + MC;
    float *zs = ys + MC;
    constexpr int PER = MC / THREADS;  // candidates per thread in the transposition
    static_assert(MC % THREADS == 0 && MC % NN_CH == 0, "chunk must tile evenly");
    __shared__ __align__(8) uint64_t mbar;
    __shared__ float warp_part[THREADS / 32];
    __shared__ int last_flag;

    const int tid = threadIdx.x;
    const int tiles0 = a.b * a.tiles[0];
    const int dir = (int)blockIdx.x >= tiles0 ? 1 : 0;
    const int t = dir ? (int)blockIdx.x - tiles0 : (int)blockIdx.x;
    const int tiles_d = dir ? a.tiles[1] : a.tiles[0];
    const int cloud = t / tiles_d;
    const int qtile = t - cloud * tiles_d;
    const int nq = dir ? a.npts[1] : a.npts[0], nc = dir ? a.npts[0] : a.npts[1];
    const float *__restrict__ Q = (dir ? a.set[1] : a.set[0]) + (size_t)cloud * nq * 3;
    const float *__restrict__ C = (dir ? a.set[0] : a.set[1]) + (size_t)cloud * nc * 3;
    float *__restrict__ out_d = (dir ? a.dist[1] : a.dist[0]) + (size_t)cloud * nq;
    int *__restrict__ out_i = (dir ? a.idx[1] : a.idx[0]) + (size_t)cloud * nq;

    if (tid == 0) mbar_init(&mbar, 1);

    // queries -> registers
    f32x2 qx[RQ], qy[RQ], qz[RQ];
    float qxs[RQ], qys[RQ], qzs[RQ];
    float best[RQ];
    int bsub[RQ];
#pragma unroll
    for (int r = 0; r < RQ; ++r) {
        int q = qtile * (THREADS * RQ) + r * THREADS + tid;
        float x = 0.f, y = 0.f, z = 0.f;
        if (q < nq) {
            x = __ldg(Q + (size_t)q * 3 + 0);
            y = __ldg(Q + (size_t)q * 3 + 1);
            z = __ldg(Q + (size_t)q * 3 + 2);
        }
        qxs[r] = x, qys[r] = y, qzs[r] = z;
        qx[r] = pack2(x, x), qy[r] = pack2(y, y), qz[r] = pack2(z, z);
        best[r] = __int_as_float(0x7f800000);
        bsub[r] = 0;
    }
    __syncthreads();  // mbarrier init visible

    uint32_t phase = 0;
    for (int c0 = 0; c0 < nc; c0 += MC) {
        const int cnt = min(MC, nc - c0);
        const float *src = C + (size_t)c0 * 3;
        const bool tma_ok = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
        const uint32_t bulk_bytes = tma_ok ? ((uint32_t)(cnt * 12) & ~15u) : 0u;
        if (bulk_bytes) {
            if (tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(&mbar, bulk_bytes);
                bulk_g2s(stage, src, bulk_bytes, &mbar);
            }
        }
        // tail (or everything, when the source is not 16-byte aligned): plain coalesced loads
        for (int i = (int)(bulk_bytes / 4) + tid; i < cnt * 3; i += THREADS) stage[i] = __ldg(src + i);
        if (bulk_bytes) {
            mbar_wait(&mbar, phase);
            phase ^= 1;
        }
        __syncthreads();
        // AoS -> SoA in place; pad the last sub-chunk with +inf coordinates (distance +inf, never a minimum)
        const int padded = (cnt + NN_CH - 1) / NN_CH * NN_CH;
        {
            float tx[PER], ty[PER], tz[PER];
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const int i = tid + j * THREADS;
                float x = __int_as_float(0x7f800000), y = x, z = x;
                if (i < cnt) x = stage[i * 3 + 0], y = stage[i * 3 + 1], z = stage[i * 3 + 2];
                tx[j] = x, ty[j] = y, tz[j] = z;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const int i = tid + j * THREADS;
                if (i < padded) xs[i] = tx[j], ys[i] = ty[j], zs[i] = tz[j];
            }
        }
        __syncthreads();

        const int nsub = padded / NN_CH;
        const int sub0 = c0 / NN_CH;
        for (int s = 0; s < nsub; ++s) {
            const ulonglong2 *x4 = reinterpret_cast<const ulonglong2 *>(xs + s * NN_CH);
            const ulonglong2 *y4 = reinterpret_cast<const ulonglong2 *>(ys + s * NN_CH);
            const ulonglong2 *z4 = reinterpret_cast<const ulonglong2 *>(zs + s * NN_CH);
            float cur[RQ];
#pragma unroll
            for (int r = 0; r < RQ; ++r) cur[r] = best[r];
#pragma unroll
            for (int v = 0; v < NN_CH / 4; ++v) {
                const ulonglong2 cx = x4[v], cy = y4[v], cz = z4[v];
#pragma unroll
                for (int r = 0; r < RQ; ++r) {
                    f32x2 d01 = sqdist_exact2(qx[r], qy[r], qz[r], cx.x, cy.x, cz.x);
                    f32x2 d23 = sqdist_exact2(qx[r], qy[r], qz[r], cx.y, cy.y, cz.y);
                    float d0, d1, d2, d3;
                    unpack2(d01, d0, d1);
                    unpack2(d23, d2, d3);
                    cur[r] = min3(cur[r], d0, d1);
                    cur[r] = min3(cur[r], d2, d3);
                }
            }
#pragma unroll
            for (int r = 0; r < RQ; ++r) {
                if (cur[r] < best[r]) {  // strict: an equal minimum in a later sub-chunk never wins
                    best[r] = cur[r];
                    bsub[r] = sub0 + s;
                }
            }
        }
        __syncthreads();  // everyone done with xs/ys/zs and stage before the next chunk lands
    }

    // argmin recovery + store
    float lsum = 0.f;
#pragma unroll
    for (int r = 0; r < RQ; ++r) {
        int q = qtile * (THREADS * RQ) + r * THREADS + tid;
        if (q < nq) {
            int base = bsub[r] * NN_CH;
            int found = base;
#pragma unroll 4
            for (int k = NN_CH - 1; k >= 0; --k) {
                int c = base + k;
                if (c < nc) {
                    float d = sqdist_exact(qxs[r], qys[r], qzs[r], __ldg(C + (size_t)c * 3 + 0),
                                           __ldg(C + (size_t)c * 3 + 1), __ldg(C + (size_t)c * 3 + 2));
                    if (d == best[r]) found = c;
                }
            }
            out_d[q] = best[r];
            out_i[q] = found;
            lsum += best[r];
        }
    }

    if (a.loss != nullptr) {
        // deterministic: fixed shuffle tree, warps in ascending order, tiles in ascending order
        float w = warp_sum(lsum);
        if ((tid & 31) == 0) warp_part[tid >> 5] = w;
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < THREADS / 32; ++i) s += warp_part[i];
            a.partial[blockIdx.x] = s;
            __threadfence();
            unsigned int prev = atomicAdd(a.counter, 1u);
            last_flag = (prev == gridDim.x - 1);
        }
        __syncthreads();
        if (last_flag) {
            __threadfence();
            float s = 0.f;
            for (int i = tid; i < (int)gridDim.x; i += THREADS) s += __ldcg(a.partial + i);
            s = warp_sum(s);
            if ((tid & 31) == 0) warp_part[tid >> 5] = s;
            __syncthreads();
            if (tid == 0) {
                float tot = 0.f;
#pragma unroll
                for (int i = 0; i < THREADS / 32; ++i) tot += warp_part[i];
                a.loss[0] = tot;
                *a.counter = 0u;  // restore the workspace invariant
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Backward.  One CTA per (cloud, side).  "own" = the set whose gradient this CTA writes,
// "other" = the opposite set.
//   grad_own[i] = 2 g_own[i] (P_i - O_{idx_own[i]})  +  sum_{k : idx_other[k]==i} -(2 g_other[k] (O_k - P_i))
// (nndistance.cu:143-151).  The second term needs the inverse of idx_other: a stable counting
// sort of k by key idx_other[k] in shared memory (integer atomics for the histogram only,
// one warp places the entries in ascending k with __match_any_sync ranks), then each thread
// gathers its bucket in ascending k -> no float atomics, bitwise reproducible.
// ------------------------------------------------------------------------------------------
struct NNGradArgs {
    const float *set[2];
    const int *idx[2];
    const float *gdist[2];  // per-point upstream grads, or (scalar_grad) one device float each
    float *grad[2];
    int npts[2];
    int b;
    int scalar_grad;
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS) nn_grad_kernel(const NNGradArgs a, const int nseg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int side = blockIdx.x & 1;
    const int cloud = blockIdx.x >> 1;
    const int np = side ? a.npts[1] : a.npts[0], no = side ? a.npts[0] : a.npts[1];
    int *keys = reinterpret_cast<int *>(smem_raw);  // [no]  idx_other staged (one coalesced pass over HBM/L2)
    int *perm = keys + no;                           // [no]  k sorted by key, stable
    int *start = perm + no;                          // [np+1] bucket starts
    int *cnt = start + (np + 1);                     // [nseg][np] per-segment counts -> per-segment cursors
    __shared__ int scan_part[THREADS];
    __shared__ int scan_total;

    const float *__restrict__ P = (side ? a.set[1] : a.set[0]) + (size_t)cloud * np * 3;
    const float *__restrict__ O = (side ? a.set[0] : a.set[1]) + (size_t)cloud * no * 3;
    const int *__restrict__ idx_own = (side ? a.idx[1] : a.idx[0]) + (size_t)cloud * np;
    const int *__restrict__ idx_oth = (side ? a.idx[0] : a.idx[1]) + (size_t)cloud * no;
    const float *__restrict__ g_own = (side ? a.gdist[1] : a.gdist[0]) + (a.scalar_grad ? 0 : (size_t)cloud * np);
    const float *__restrict__ g_oth = (side ? a.gdist[0] : a.gdist[1]) + (a.scalar_grad ? 0 : (size_t)cloud * no);
    float *__restrict__ G = (side ? a.grad[1] : a.grad[0]) + (size_t)cloud * np * 3;
    const int tid = threadIdx.x;
    // k-range of placement segment s: [s*seg, min(no,(s+1)*seg)), seg a multiple of 32
    const int seg = ((no + nseg - 1) / nseg + 31) & ~31;

    // Issue the loads of the "own" term first: their latency hides behind the sort phases below.
    constexpr int PRE = 2;  // points per thread kept in registers (covers np <= PRE*THREADS)
    const float gs_own = a.scalar_grad ? __ldg(g_own) : 0.f;
    const float gs_oth = a.scalar_grad ? __ldg(g_oth) : 0.f;
    float pre_p[PRE][3], pre_o[PRE][3], pre_g[PRE];
#pragma unroll
    for (int u = 0; u < PRE; ++u) {
        const int i = tid + u * THREADS;
        if (i < np) {
            const int j2 = min(max(__ldg(idx_own + i), 0), no - 1);
#pragma unroll
            for (int c = 0; c < 3; ++c) pre_p[u][c] = __ldg(P + (size_t)i * 3 + c), pre_o[u][c] = __ldg(O + (size_t)j2 * 3 + c);
            pre_g[u] = (a.scalar_grad ? gs_own : __ldg(g_own + i)) * 2.f;
        }
    }

    for (int i = tid; i < nseg * np; i += THREADS) cnt[i] = 0;
    for (int k = tid; k < no; k += THREADS) {
        int key = __ldg(idx_oth + k);
        keys[k] = ((unsigned)key < (unsigned)np) ? key : -1;
    }
    __syncthreads();
    for (int k = tid; k < no; k += THREADS) {
        int key = keys[k];
        if (key >= 0) atomicAdd(&cnt[(k / seg) * np + key], 1);  // integer atomics: order-independent result
    }
    __syncthreads();
    // per key: exclusive prefix over segments (in place) and the bucket size (into start[])
    for (int i = tid; i < np; i += THREADS) {
        int run = 0;
        for (int s = 0; s < nseg; ++s) {
            int c = cnt[s * np + i];
            cnt[s * np + i] = run;
            run += c;
        }
        start[i] = run;
    }
    __syncthreads();
    // exclusive scan of start[0..np): contiguous slice per thread + one-warp scan of the partials
    const int per = (np + THREADS - 1) / THREADS;
    const int lo = min(np, tid * per), hi = min(np, lo + per);
    int local = 0;
    for (int i = lo; i < hi; ++i) local += start[i];
    scan_part[tid] = local;
    __syncthreads();
    if (tid < 32) {
        int carry = 0;
        for (int base = 0; base < THREADS; base += 32) {
            int v = scan_part[base + tid];
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int n = __shfl_up_sync(0xffffffffu, inc, o);
                if (tid >= o) inc += n;
            }
            scan_part[base + tid] = carry + inc - v;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (tid == 0) scan_total = carry;
    }
    __syncthreads();
    {
        int run = scan_part[tid];
        for (int i = lo; i < hi; ++i) {
            int c = start[i];
            start[i] = run;
            run += c;
        }
    }
    if (tid == 0) start[np] = scan_total;
    __syncthreads();
    // stable placement: warp w owns segment w and walks it in ascending k, 32 at a time
    const int warp = tid >> 5, lane = tid & 31;
    if (warp < nseg) {
        int *cur = cnt + warp * np;
        const int k_end = min(no, (warp + 1) * seg);
        for (int k0 = warp * seg; k0 < k_end; k0 += 32) {
            int k = k0 + lane;
            int key = (k < k_end) ? keys[k] : -1;
            bool ok = key >= 0;
            unsigned act = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                unsigned peers = __match_any_sync(act, key);
                int rank = __popc(peers & ((1u << lane) - 1u));
                int basep = cur[key];
                __syncwarp(act);
                perm[start[key] + basep + rank] = k;
                if (rank == __popc(peers) - 1) cur[key] = basep + rank + 1;
            }
            __syncwarp();
        }
    }
    __syncthreads();

    // Gather.  Buckets up to GRAD_COOP entries are walked by their owner in ascending k; larger ones (skewed assignments)
    // by the whole warp: lane l takes entries l, l+32, ... and the partial sums are folded by a fixed shuffle tree.
    constexpr int GRAD_COOP = 32;
    const int lane_g = tid & 31;
    for (int base = 0; base < np; base += THREADS) {  // block-uniform trip count: the warp-cooperative part needs all lanes
        const int i = base + tid;
        const bool valid = i < np;
        const int u = base / THREADS;
        float px = 0.f, py = 0.f, pz = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
        int pb = 0, pe = 0;
        if (valid) {
            float ox, oy, oz, g;
            static_assert(PRE == 2, "the preloaded points are selected with static register indices");
            if (u < PRE) {
                px = u ? pre_p[1][0] : pre_p[0][0], py = u ? pre_p[1][1] : pre_p[0][1], pz = u ? pre_p[1][2] : pre_p[0][2];
                ox = u ? pre_o[1][0] : pre_o[0][0], oy = u ? pre_o[1][1] : pre_o[0][1], oz = u ? pre_o[1][2] : pre_o[0][2];
                g = u ? pre_g[1] : pre_g[0];
            } else {
                const int j2 = min(max(__ldg(idx_own + i), 0), no - 1);
                px = __ldg(P + (size_t)i * 3 + 0), py = __ldg(P + (size_t)i * 3 + 1), pz = __ldg(P + (size_t)i * 3 + 2);
                ox = __ldg(O + (size_t)j2 * 3 + 0), oy = __ldg(O + (size_t)j2 * 3 + 1), oz = __ldg(O + (size_t)j2 * 3 + 2);
                g = (a.scalar_grad ? gs_own : __ldg(g_own + i)) * 2.f;
            }
            ax = g * (px - ox), ay = g * (py - oy), az = g * (pz - oz);
            pb = start[i], pe = start[i + 1];
        }
        const bool big = valid && (pe - pb) > GRAD_COOP;
        if (valid && !big) {
            for (int p = pb; p < pe; ++p) {  // ascending k: fixed summation order
                const int k = perm[p];
                const float gk = (a.scalar_grad ? gs_oth : __ldg(g_oth + k)) * 2.f;
                ax += -(gk * (__ldg(O + (size_t)k * 3 + 0) - px));
                ay += -(gk * (__ldg(O + (size_t)k * 3 + 1) - py));
                az += -(gk * (__ldg(O + (size_t)k * 3 + 2) - pz));
            }
        }
        unsigned todo = __ballot_sync(0xffffffffu, big);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int b0 = __shfl_sync(0xffffffffu, pb, src), b1 = __shfl_sync(0xffffffffu, pe, src);
            const float qx = __shfl_sync(0xffffffffu, px, src), qy = __shfl_sync(0xffffffffu, py, src), qz = __shfl_sync(0xffffffffu, pz, src);
            float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll 4
            for (int p = b0 + lane_g; p < b1; p += 32) {
                const int k = perm[p];
                const float gk = (a.scalar_grad ? gs_oth : __ldg(g_oth + k)) * 2.f;
                sx += -(gk * (__ldg(O + (size_t)k * 3 + 0) - qx));
                sy += -(gk * (__ldg(O + (size_t)k * 3 + 1) - qy));
                sz += -(gk * (__ldg(O + (size_t)k * 3 + 2) - qz));
            }
            sx = warp_sum(sx), sy = warp_sum(sy), sz = warp_sum(sz);
            if (lane_g == src) ax += sx, ay += sy, az += sz;
        }
        if (valid) {
            G[(size_t)i * 3 + 0] = ax;
            G[(size_t)i * 3 + 1] = ay;
            G[(size_t)i * 3 + 2] = az;
        }
    }
}

// Large-cloud fallback (n+m > HP_NNGRAD_SMEM_POINTS): float atomics like the reference.
__global__ void nn_grad_atomic_kernel(int b, int n, const float *__restrict__ xyz1, int m,
                                      const float *__restrict__ xyz2, const float *__restrict__ gd,
                                      int scalar_grad, const int *__restrict__ idx1, float *grad1, float *grad2) {
    size_t total = (size_t)b * n;
    for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
        size_t i = o / n;
        int j2 = idx1[o];
        size_t o2 = i * m + j2;
        float g = (scalar_grad ? gd[0] : gd[o]) * 2.f;
        float dx = xyz1[o * 3 + 0] - xyz2[o2 * 3 + 0];
        float dy = xyz1[o * 3 + 1] - xyz2[o2 * 3 + 1];
        float dz = xyz1[o * 3 + 2] - xyz2[o2 * 3 + 2];
        atomicAdd(grad1 + o * 3 + 0, g * dx);
        atomicAdd(grad1 + o * 3 + 1, g * dy);
        atomicAdd(grad1 + o * 3 + 2, g * dz);
        atomicAdd(grad2 + o2 * 3 + 0, -(g * dx));
        atomicAdd(grad2 + o2 * 3 + 1, -(g * dy));
        atomicAdd(grad2 + o2 * 3 + 2, -(g * dz));
    }
}

// warp-ring forward (chamfer_ring.cu): every unordered pair once, both directions
size_t nn_ring_workspace_bytes(int b, int n, int m);
int nn_ring_forward_launch(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1, int *idx1, float *dist2,
                           int *idx2, float *loss, int *inv1, int *inv2, void *workspace, cudaStream_t stream);
bool nn_ring_inverse_supported(int n, int m);
bool nn_ring_step_supported(int n, int m);
#ifdef HP_BENCH_BUILD
int nn_ring_set_trace(void *dev_ptr);
int nn_ring_only_launch(int b, int n, const float *xyz1, int m, const float *xyz2, void *workspace, cudaStream_t stream);
#endif
int nn_ring_step_launch(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_loss, float *dist1, int *idx1,
                        float *dist2, int *idx2, float *loss, float *grad1, float *grad2, void *workspace, cudaStream_t stream);
int nn_ring_backward_launch(int b, int n, const float *xyz1, int m, const float *xyz2, const int *idx1, const int *idx2,
                            const int *inv1, const int *inv2, const float *grad_loss, const float *grad_dist1,
                            const float *grad_dist2, float *grad1, float *grad2, cudaStream_t stream);

// The first-generation ordered-pair kernel stays as the fallback of hp_nndistance when no stream-ordered workspace can be had;
// the bench library (HP_BENCH_BUILD) can select it with HP_NN_RING=0 for A/B measurements.
static bool use_ring() {
#ifdef HP_BENCH_BUILD
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("HP_NN_RING");
        v = (e && atoi(e) == 0) ? 0 : 1;
    }
    return v == 1;
#else
    return true;
#endif
}

// ---- host side ---------------------------------------------------------------------------
constexpr int FWD_MIN_QT = 32;  // smallest query tile of any variant (sizes the loss workspace)
constexpr int GRAD_THREADS = 1024;

// Forward tile variants (threads per CTA, queries per thread) of the ordered-pair kernel; the bench library can select one
// with HP_NN_VARIANT, the product build uses the best measured on B200 at B=32, N=M=2048.
static int fwd_variant() {
#ifdef HP_BENCH_BUILD
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("HP_NN_VARIANT");
        v = e ? atoi(e) : 0;
        if (v < 0 || v > 8) v = 0;
    }
    return v;
#else
    return 0;
#endif
}

template <int THREADS, int RQ, int MC>
static int nn_forward_launch_t(NNArgs a, cudaStream_t stream) {
    const int QT = THREADS * RQ;
    a.tiles[0] = (a.npts[0] + QT - 1) / QT;
    a.tiles[1] = (a.npts[1] + QT - 1) / QT;
    long long grid = (long long)a.b * (a.tiles[0] + a.tiles[1]);
    HP_REQUIRE(grid <= 0x7fffffffLL, "hp_nndistance: grid too large (%lld tiles)", grid);
    auto kern = nn_fwd_kernel<THREADS, RQ, MC>;
    const size_t smem = (size_t)MC * 3 * sizeof(float);
    static SmemAttrCache fwd_attr;
    HP_CUDA(ensure_dynamic_smem(kern, smem, fwd_attr));
    kern<<<(unsigned)grid, THREADS, smem, stream>>>(a);
    HP_LAUNCH_CHECK("nn_fwd_kernel");
    return HP_OK;
}

static int nn_forward_launch(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1, int *idx1,
                             float *dist2, int *idx2, float *loss, void *workspace, cudaStream_t stream) {
    NNArgs a;
    a.set[0] = xyz1, a.set[1] = xyz2;
    a.dist[0] = dist1, a.dist[1] = dist2;
    a.idx[0] = idx1, a.idx[1] = idx2;
    a.npts[0] = n, a.npts[1] = m;
    a.b = b;
    a.loss = loss;
    a.counter = reinterpret_cast<unsigned int *>(workspace);
    a.partial = workspace ? reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(workspace) + 8) : nullptr;
    switch (fwd_variant()) {
#ifdef HP_BENCH_BUILD
        case 1: return nn_forward_launch_t<128, 2, 2048>(a, stream);
        case 2: return nn_forward_launch_t<64, 2, 1024>(a, stream);
        case 3: return nn_forward_launch_t<256, 1, 2048>(a, stream);
        case 4: return nn_forward_launch_t<64, 1, 1024>(a, stream);
        case 5: return nn_forward_launch_t<64, 4, 1024>(a, stream);
        case 6: return nn_forward_launch_t<128, 4, 2048>(a, stream);
        case 7: return nn_forward_launch_t<32, 4, 512>(a, stream);
        case 8: return nn_forward_launch_t<32, 8, 512>(a, stream);
#endif
        default: return nn_forward_launch_t<128, 1, 2048>(a, stream);
    }
}

static int nn_backward_launch(int b, int n, const float *xyz1, int m, const float *xyz2, const float *g1, const int *idx1,
                              const float *g2, const int *idx2, float *grad1, float *grad2, int scalar_grad,
                              cudaStream_t stream) {
    // shared memory (ints) of one (cloud, side) CTA: keys[no] perm[no] start[np+1] cnt[nseg][np] -- sized for the worse of the
    // two sides, with as many placement segments (<= one per warp) as fit; clouds that do not fit even one segment
    // (n + m above HP_NNGRAD_SMEM_POINTS, or very unbalanced pairs such as 16384 x 2048) take the atomicAdd kernel below.
    const size_t budget = (size_t)200 * 1024 / sizeof(int);
    auto side_ints = [](size_t np, size_t no, size_t nseg) { return 2 * no + np + 1 + nseg * np; };
    int nseg = 0;
    if ((long long)n + m <= HP_NNGRAD_SMEM_POINTS) {
        for (int s = GRAD_THREADS / 32; s >= 1; --s) {
            if (side_ints(n, m, s) <= budget && side_ints(m, n, s) <= budget) {
                nseg = s;
                break;
            }
        }
    }
    if (nseg >= 1) {
        NNGradArgs a;
        a.set[0] = xyz1, a.set[1] = xyz2;
        a.idx[0] = idx1, a.idx[1] = idx2;
        a.gdist[0] = g1, a.gdist[1] = g2;
        a.grad[0] = grad1, a.grad[1] = grad2;
        a.npts[0] = n, a.npts[1] = m;
        a.b = b;
        a.scalar_grad = scalar_grad;
        const size_t ints = side_ints(n, m, nseg) > side_ints(m, n, nseg) ? side_ints(n, m, nseg) : side_ints(m, n, nseg);
        const size_t smem = ints * sizeof(int);
        auto kern = nn_grad_kernel<GRAD_THREADS>;
        static SmemAttrCache grad_attr;
        if (smem > 48 * 1024) HP_CUDA(ensure_dynamic_smem(kern, smem, grad_attr));
        kern<<<2 * b, GRAD_THREADS, smem, stream>>>(a, nseg);
        HP_LAUNCH_CHECK("nn_grad_kernel");
        return HP_OK;
    }
    HP_CUDA(cudaMemsetAsync(grad1, 0, (size_t)b * n * 3 * sizeof(float), stream));
    HP_CUDA(cudaMemsetAsync(grad2, 0, (size_t)b * m * 3 * sizeof(float), stream));
    int blocks = sm_count() * 8;
    nn_grad_atomic_kernel<<<blocks, 256, 0, stream>>>(b, n, xyz1, m, xyz2, g1, scalar_grad, idx1, grad1, grad2);
    HP_LAUNCH_CHECK("nn_grad_atomic_kernel(1)");
    nn_grad_atomic_kernel<<<blocks, 256, 0, stream>>>(b, m, xyz2, n, xyz1, g2, scalar_grad, idx2, grad2, grad1);
    HP_LAUNCH_CHECK("nn_grad_atomic_kernel(2)");
    return HP_OK;
}

}  // namespace hp

using namespace hp;

extern "C" int hp_nndistance(int b, int n, const float *xyz, int m, const float *xyz2, float *result, int *result_i,
                             float *result2, int *result2_i, void *stream) {
    HP_REQUIRE(b >= 0 && n >= 0 && m >= 0, "hp_nndistance: negative size (b=%d n=%d m=%d)", b, n, m);
    if (b == 0 || (n == 0 && m == 0)) return HP_OK;
    HP_REQUIRE(n > 0 && m > 0, "hp_nndistance: one point set is empty (n=%d m=%d): nearest neighbour undefined", n, m);
    HP_REQUIRE(xyz && xyz2 && result && result_i && result2 && result2_i, "hp_nndistance: null pointer");
    // The reference signature carries no workspace: take one from the stream-ordered pool (cudaMallocAsync / cudaFreeAsync:
    // no synchronisation, legal under stream capture) and run the ring kernels like hp_nndistance_ws.  If the pool refuses,
    // the workspace-free ordered-pair kernel computes the same bits at about half the speed.
    cudaStream_t st = (cudaStream_t)stream;
    if (use_ring()) {
        const size_t bytes = nn_ring_workspace_bytes(b, n, m);
        void *ws = nullptr;
        if (cudaMallocAsync(&ws, bytes, st) == cudaSuccess && ws != nullptr) {
            int rc = check_cuda(cudaMemsetAsync(ws, 0, bytes, st), "hp_nndistance: cudaMemsetAsync");
            if (rc == HP_OK)
                rc = nn_ring_forward_launch(b, n, xyz, m, xyz2, result, result_i, result2, result2_i, nullptr, nullptr, nullptr, ws, st);
            const int rf = check_cuda(cudaFreeAsync(ws, st), "hp_nndistance: cudaFreeAsync");
            return rc != HP_OK ? rc : rf;
        }
        (void)cudaGetLastError();  // pool unavailable: not an error of this call
    }
    return nn_forward_launch(b, n, xyz, m, xyz2, result, result_i, result2, result2_i, nullptr, nullptr, st);
}

extern "C" size_t hp_chamfer_workspace_bytes(int b, int n, int m) {
    if (b <= 0 || n <= 0 || m <= 0) return 16;
    const int QT = FWD_MIN_QT;
    size_t tiles = (size_t)b * ((n + QT - 1) / QT + (m + QT - 1) / QT);
    const size_t ordered = 8 + tiles * sizeof(float) + 8, ring = nn_ring_workspace_bytes(b, n, m);
    return ordered > ring ? ordered : ring;
}

extern "C" int hp_nndistance_ws(int b, int n, const float *xyz, int m, const float *xyz2, float *result, int *result_i,
                                float *result2, int *result2_i, void *workspace, size_t workspace_bytes, void *stream) {
    HP_REQUIRE(b >= 0 && n >= 0 && m >= 0, "hp_nndistance_ws: negative size (b=%d n=%d m=%d)", b, n, m);
    if (b == 0 || (n == 0 && m == 0)) return HP_OK;
    HP_REQUIRE(n > 0 && m > 0, "hp_nndistance_ws: one point set is empty (n=%d m=%d): nearest neighbour undefined", n, m);
    HP_REQUIRE(xyz && xyz2 && result && result_i && result2 && result2_i, "hp_nndistance_ws: null pointer");
    HP_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
               "hp_nndistance_ws: workspace null or not 16-byte aligned");
    if (workspace_bytes < hp_chamfer_workspace_bytes(b, n, m)) {
        set_error("hp_nndistance_ws: workspace %zu < required %zu bytes", workspace_bytes, hp_chamfer_workspace_bytes(b, n, m));
        return HP_ERR_WORKSPACE;
    }
    if (use_ring())
        return nn_ring_forward_launch(b, n, xyz, m, xyz2, result, result_i, result2, result2_i, nullptr, nullptr, nullptr,
                                      workspace, (cudaStream_t)stream);
    return nn_forward_launch(b, n, xyz, m, xyz2, result, result_i, result2, result2_i, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int hp_chamfer_forward(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1, int *idx1,
                                  float *dist2, int *idx2, float *loss, void *workspace, size_t workspace_bytes,
                                  void *stream) {
    HP_REQUIRE(b >= 0 && n >= 0 && m >= 0, "hp_chamfer_forward: negative size (b=%d n=%d m=%d)", b, n, m);
    HP_REQUIRE(loss != nullptr, "hp_chamfer_forward: loss pointer is null");
    if (b == 0 || (n == 0 && m == 0)) {
        HP_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream));
        return HP_OK;
    }
    HP_REQUIRE(n > 0 && m > 0, "hp_chamfer_forward: one point set is empty (n=%d m=%d)", n, m);
    HP_REQUIRE(xyz1 && xyz2 && dist1 && idx1 && dist2 && idx2, "hp_chamfer_forward: null pointer");
    HP_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
               "hp_chamfer_forward: workspace null or not 16-byte aligned");
    if (workspace_bytes < hp_chamfer_workspace_bytes(b, n, m)) {
        set_error("hp_chamfer_forward: workspace %zu < required %zu bytes", workspace_bytes,
                  hp_chamfer_workspace_bytes(b, n, m));
        return HP_ERR_WORKSPACE;
    }
    if (use_ring())
        return nn_ring_forward_launch(b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, loss, nullptr, nullptr, workspace,
                                      (cudaStream_t)stream);
    return nn_forward_launch(b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, loss, workspace, (cudaStream_t)stream);
}

extern "C" size_t hp_chamfer_inverse_ints(int b, int n, int m, int which) {
    if (b <= 0 || n <= 0 || m <= 0 || !use_ring() || !nn_ring_inverse_supported(n, m)) return 0;
    return which == 2 ? (size_t)b * ((size_t)m + 2 * (size_t)n) : (size_t)b * ((size_t)n + 2 * (size_t)m);
}

extern "C" int hp_chamfer_forward_inv(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1, int *idx1,
                                      float *dist2, int *idx2, float *loss, int *inv1, int *inv2, void *workspace,
                                      size_t workspace_bytes, void *stream) {
    HP_REQUIRE(b > 0 && n > 0 && m > 0, "hp_chamfer_forward_inv: sizes must be positive (b=%d n=%d m=%d)", b, n, m);
    HP_REQUIRE(xyz1 && xyz2 && dist1 && idx1 && dist2 && idx2 && inv1 && inv2, "hp_chamfer_forward_inv: null pointer");
    HP_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
               "hp_chamfer_forward_inv: workspace null or not 16-byte aligned");
    if (workspace_bytes < hp_chamfer_workspace_bytes(b, n, m)) {
        set_error("hp_chamfer_forward_inv: workspace %zu < required %zu bytes", workspace_bytes, hp_chamfer_workspace_bytes(b, n, m));
        return HP_ERR_WORKSPACE;
    }
    if (!use_ring() || !nn_ring_inverse_supported(n, m)) {
        set_error("hp_chamfer_forward_inv: inverse maps unavailable for n=%d m=%d (use hp_chamfer_forward + hp_chamfer_backward)", n, m);
        return HP_ERR_UNSUPPORTED;
    }
    return nn_ring_forward_launch(b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, loss, inv1, inv2, workspace, (cudaStream_t)stream);
}

extern "C" int hp_chamfer_backward_inv(int b, int n, const float *xyz1, int m, const float *xyz2, const int *idx1, const int *idx2,
                                       const int *inv1, const int *inv2, const float *grad_loss, float *grad_xyz1,
                                       float *grad_xyz2, void *stream) {
    HP_REQUIRE(b > 0 && n > 0 && m > 0, "hp_chamfer_backward_inv: sizes must be positive (b=%d n=%d m=%d)", b, n, m);
    HP_REQUIRE(xyz1 && xyz2 && idx1 && idx2 && inv1 && inv2 && grad_loss && grad_xyz1 && grad_xyz2,
               "hp_chamfer_backward_inv: null pointer");
    return nn_ring_backward_launch(b, n, xyz1, m, xyz2, idx1, idx2, inv1, inv2, grad_loss, nullptr, nullptr, grad_xyz1,
                                   grad_xyz2, (cudaStream_t)stream);
}

#ifdef HP_BENCH_BUILD
extern "C" HP_API int hp_measure_chamfer_ring_only(int b, int n, const float *xyz1, int m, const float *xyz2, void *workspace,
                                                   size_t workspace_bytes, void *stream) {
    HP_REQUIRE(b > 0 && n > 0 && m > 0 && xyz1 && xyz2, "hp_measure_chamfer_ring_only: bad arguments");
    HP_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0 &&
                   workspace_bytes >= hp_chamfer_workspace_bytes(b, n, m),
               "hp_measure_chamfer_ring_only: workspace null, misaligned or too small");
    return nn_ring_only_launch(b, n, xyz1, m, xyz2, workspace, (cudaStream_t)stream);
}
extern "C" HP_API int hp_measure_set_trace(void *device_u64_buffer) { return nn_ring_set_trace(device_u64_buffer); }
#endif

extern "C" int hp_chamfer_step_supported(int b, int n, int m) {
    return (b > 0 && use_ring() && nn_ring_step_supported(n, m)) ? 1 : 0;
}

extern "C" int hp_chamfer_step(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_loss, float *dist1,
                               int *idx1, float *dist2, int *idx2, float *loss, float *grad_xyz1, float *grad_xyz2,
                               void *workspace, size_t workspace_bytes, void *stream) {
    HP_REQUIRE(b > 0 && n > 0 && m > 0, "hp_chamfer_step: sizes must be positive (b=%d n=%d m=%d)", b, n, m);
    HP_REQUIRE(xyz1 && xyz2 && grad_loss && dist1 && idx1 && dist2 && idx2 && loss && grad_xyz1 && grad_xyz2,
               "hp_chamfer_step: null pointer");
    HP_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
               "hp_chamfer_step: workspace null or not 16-byte aligned");
    if (workspace_bytes < hp_chamfer_workspace_bytes(b, n, m)) {
        set_error("hp_chamfer_step: workspace %zu < required %zu bytes", workspace_bytes, hp_chamfer_workspace_bytes(b, n, m));
        return HP_ERR_WORKSPACE;
    }
    if (!hp_chamfer_step_supported(b, n, m)) {
        set_error("hp_chamfer_step: fused step unavailable for n=%d m=%d (use hp_chamfer_forward_inv + hp_chamfer_backward_inv)", n, m);
        return HP_ERR_UNSUPPORTED;
    }
    return nn_ring_step_launch(b, n, xyz1, m, xyz2, grad_loss, dist1, idx1, dist2, idx2, loss, grad_xyz1, grad_xyz2, workspace,
                               (cudaStream_t)stream);
}

extern "C" int hp_nndistancegrad_inv(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_dist1,
                                     const int *idx1, const float *grad_dist2, const int *idx2, const int *inv1,
                                     const int *inv2, float *grad_xyz1, float *grad_xyz2, void *stream) {
    HP_REQUIRE(b > 0 && n > 0 && m > 0, "hp_nndistancegrad_inv: sizes must be positive (b=%d n=%d m=%d)", b, n, m);
    HP_REQUIRE(xyz1 && xyz2 && grad_dist1 && idx1 && grad_dist2 && idx2 && inv1 && inv2 && grad_xyz1 && grad_xyz2,
               "hp_nndistancegrad_inv: null pointer");
    return nn_ring_backward_launch(b, n, xyz1, m, xyz2, idx1, idx2, inv1, inv2, nullptr, grad_dist1, grad_dist2, grad_xyz1,
                                   grad_xyz2, (cudaStream_t)stream);
}

extern "C" int hp_nndistancegrad(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_dist1,
                                 const int *idx1, const float *grad_dist2, const int *idx2, float *grad_xyz1,
                                 float *grad_xyz2, void *stream) {
    HP_REQUIRE(b >= 0 && n >= 0 && m >= 0, "hp_nndistancegrad: negative size (b=%d n=%d m=%d)", b, n, m);
    if (b == 0 || (n == 0 && m == 0)) return HP_OK;
    HP_REQUIRE(n > 0 && m > 0, "hp_nndistancegrad: one point set is empty (n=%d m=%d)", n, m);
    HP_REQUIRE(xyz1 && xyz2 && grad_dist1 && idx1 && grad_dist2 && idx2 && grad_xyz1 && grad_xyz2,
               "hp_nndistancegrad: null pointer");
    return nn_backward_launch(b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2, 0,
                              (cudaStream_t)stream);
}

extern "C" int hp_chamfer_backward(int b, int n, const float *xyz1, int m, const float *xyz2, const int *idx1,
                                   const int *idx2, const float *grad_loss, float *grad_xyz1, float *grad_xyz2,
                                   void *stream) {
    HP_REQUIRE(b >= 0 && n >= 0 && m >= 0, "hp_chamfer_backward: negative size (b=%d n=%d m=%d)", b, n, m);
    if (b == 0 || (n == 0 && m == 0)) return HP_OK;
    HP_REQUIRE(n > 0 && m > 0, "hp_chamfer_backward: one point set is empty (n=%d m=%d)", n, m);
    HP_REQUIRE(xyz1 && xyz2 && idx1 && idx2 && grad_loss && grad_xyz1 && grad_xyz2, "hp_chamfer_backward: null pointer");
    return nn_backward_launch(b, n, xyz1, m, xyz2, grad_loss, idx1, grad_loss, idx2, grad_xyz1, grad_xyz2, 1,
                              (cudaStream_t)stream);
}
