// chamfer_ring.cu -- nearest-neighbour forward (values AND indices, both directions) in "warp ring" form.
//
// The first-generation kernel (chamfer.cu: nn_fwd_kernel) evaluates every ORDERED (query, candidate) pair:
// 2*B*N*M distance evaluations.  d(a_i, b_j) is bit-symmetric under the library's one association
// fma(dz,dz,fma(dx,dx,dy*dy)), so here every UNORDERED pair is evaluated once and feeds both directions:
//
//   * a lane keeps 8 points of set 1 ("rows") in registers; the 128 columns of a round sit in shared memory as 32 packed
//     groups of 4.  32 times per round a lane evaluates its 8 rows against one group (FADD2/FMUL2/FFMA2, 96 packed ops),
//     folds the 8x4 distances into its 8 row minima and into the group's 4 column minima, and the group's STATE (minima +
//     rotation of last improvement) is handed to the next lane through shared memory (LDS.128 / STS.128 + __syncwarp).
//     The loop is bound by instruction ISSUE: a packed fp32x2 op holds the issue port for two cycles, everything else for
//     one -- 2*96 + 88 = 280 slots per rotation (hp_measure_peak kinds 7-12).
//   * indices: per row / per column only the ROTATION in which the minimum last improved is tracked
//     (FSETP+SEL per row per rotation: 24 of the 280 issue slots); afterwards the 4 columns
//     (8 rows) met in that rotation are re-evaluated with bit-identical arithmetic and the lowest index with
//     d == min is taken.
//   * ties: the reference keeps the LOWEST index among equal distances (nndistance.cu:32-64,117-125).  A lane
//     meets column groups in ascending index order except for ONE wrap (groups 31-L+t mod 32), so a plain
//     "strict <" would favour the pre-wrap group.  At the wrap the running minimum is bumped by one ulp (as
//     integer bits: d >= 0), which makes "strict <" behave like "<=" for the post-wrap (lower-index) groups
//     only; the bump is removed afterwards if nothing replaced it.  Same for columns, which wrap when they
//     pass from lane 31 to lane 0.  Result: bit-exact distances and reference-exact indices.
//   * a CTA (4 warps) owns 1024 rows x R*128 columns of one cloud; per-CTA (distance,index) candidates are merged with
//     64-bit atomicMax on ~(float bits << 32 | index) keys (integer order == (distance, index) order) in a zero-restored
//     workspace.  The keys are consumed either by nn_ring_finish_kernel (training step: distances, indices, loss, inverse
//     index maps in shared memory and BOTH gradients in one kernel, launched programmatically dependent) or by
//     nn_ring_unpack_kernel (+ nn_grad_gather_kernel later), when the upstream gradient is not known yet.
#include <stdlib.h>

#include "common.cuh"

namespace hp {

constexpr int RF_RQ = 8, RF_RC = 4;
constexpr int RF_WROWS = 32 * RF_RQ;            // 256 rows per warp
#ifndef HP_RING_WARPS
#define HP_RING_WARPS 4
#endif
constexpr int RF_DEFAULT_WARPS = HP_RING_WARPS; // warps per CTA of the product build: a CTA owns RF_DEFAULT_WARPS x 256 rows
constexpr int RF_COLS = 32 * RF_RC;             // 128 columns per round
constexpr float RF_PAD_ROW = 1.0e18f;           // padding points: finite, farther than any real pair
constexpr float RF_PAD_COL = -1.0e18f;
constexpr int RF_MERGE_THREADS = 1024;
constexpr int RF_MAX_PER_THREAD = 32;           // 32768 / RF_MERGE_THREADS: largest cloud of the in-kernel inverse

typedef unsigned long long u64;

#ifdef HP_BENCH_BUILD
// bench library only: per-CTA timeline (globaltimer) of the ring and tail kernels, see hp_measure_set_trace
static unsigned long long *g_trace_host = nullptr;  // handed to the kernels through their argument struct
__device__ __forceinline__ void trace_mark(unsigned long long *trace, size_t slot) {
    if (trace != nullptr) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        trace[slot] = t;
    }
}
#define HP_TRACE(slot) do { if (threadIdx.x == 0) trace_mark(a.trace, slot); } while (0)
#define HP_TRACE_T(thread, slot) do { if (threadIdx.x == (thread)) trace_mark(a.trace, slot); } while (0)
#else
#define HP_TRACE(slot) do { } while (0)
#define HP_TRACE_T(thread, slot) do { } while (0)
#endif

struct RingNNArgs {
    const float *set1, *set2;  // [b,n,3], [b,m,3]
    int b, n, m;
    int rowchunks, colchunks, R;
    u64 *rowkey;               // [b][n]  ~key of the best candidate so far (0 = none): zero on entry, zero on exit
    u64 *colkey;               // [b][m]
    float *dist1, *dist2;
    int *idx1, *idx2;
    float *loss;               // nullptr: no fused loss
    float *losspart;           // loss partials: [2b] per (cloud, direction) (unpack kernel) or per 256-source chunk (tail kernel)
    unsigned int *counters;    // [1] loss ticket; zero on entry, zero on exit
    unsigned int *ticket;      // [b] ring CTAs of the cloud that have merged their keys; zero on entry, zero on exit
    unsigned int *done;        // [b] tail CTAs of the cloud that have read the keys; zero on entry, zero on exit
    unsigned long long *trace; // bench library only: per-CTA timeline (nullptr = off)
    int stage_sources;         // tail kernel: 1 = the direction's source points are staged in shared memory, 0 = read through L1
    int tickets;               // 1: ring CTAs bump the per-cloud ticket (tail CTAs wait per cloud); 0: the tail waits for the whole grid
    // optional inverse index maps for the atomic-free backward (nullptr: not produced).
    //   inv1 [b][n + 2m]: perm1[n] = row indices i sorted by (idx1[i], i); then begin1[m], end1[m]: bucket of column k
    //   inv2 [b][m + 2n]: perm2[m] = column indices k sorted by (idx2[k], k); then begin2[n], end2[n]: bucket of row i
    int *inv1, *inv2;
    int sort_n, sort_shift;    // power-of-two sort size and bit position of the key in the composite (key << shift | pos)
    int inv_fast;              // 1: counting sort with radix fallback in shared memory, 0: bitonic only
};

// Key merge: a REDUCTION (no return value -> SASS RED, fire and forget).  Written in PTX on purpose: with a fence later in the
// kernel nvcc turns atomicMax() with an unused result into ATOM, and a warp cannot exit while an ATOM's return is outstanding --
// every CTA then lingers for an L2 round trip at its end (measured: the ring kernel 48 -> 52 us).
__device__ __forceinline__ void red_max_u64(u64 *addr, u64 v) {
    asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void red_add_u32(unsigned int *addr, unsigned int v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ float bump_up(float v) { return __int_as_float(__float_as_int(v) + 1); }
__device__ __forceinline__ float bump_down(float v) { return __int_as_float(__float_as_int(v) - 1); }

// Staging of `cnt` points (AoS) from global to shared memory: bulk-TMA for the 16-byte aligned body (rf_bulk_bytes, issued
// by one thread under an mbarrier), plain loads for the rest and padding up to `total` points with `pad` (rf_stage_tail).
__device__ __forceinline__ uint32_t rf_bulk_bytes(const float *src, int cnt) {
    return (reinterpret_cast<uintptr_t>(src) & 15) == 0 ? ((uint32_t)(cnt * 12) & ~15u) : 0u;
}
__device__ __forceinline__ void rf_stage_tail(float *dst, const float *src, int cnt, int total, float pad, uint32_t bulk_bytes,
                                              int tid, int nthreads) {
    for (int i = (int)(bulk_bytes / 4) + tid; i < cnt * 3; i += nthreads) dst[i] = __ldg(src + i);
    for (int i = cnt * 3 + tid; i < total * 3; i += nthreads) dst[i] = pad;
}

// A column group = 4 consecutive columns stored as 12 floats [x0 x1 y0 y1 | z0 z1 x2 x3 | y2 y3 z2 z3]: three LDS.128
// deliver the six fp32x2 operands of the packed distance evaluation in aligned register pairs, no moves.
struct RFGroup {
    f32x2 x01, y01, z01, x23, y23, z23;
};
__device__ __forceinline__ RFGroup rf_load_group(const float *cols_p, int g) {
    const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(cols_p + g * 12);
    const ulonglong2 a = p[0], b = p[1], c = p[2];
    RFGroup r;
    r.x01 = a.x, r.y01 = a.y, r.z01 = b.x, r.x23 = b.y, r.y23 = c.x, r.z23 = c.y;
    return r;
}

// DBG (bench library only, HP_BENCH_BUILD): timing experiments, bit 0 = two rotations instead of 32, bit 1 = no rescans,
// bit 2 = no ticket arrival
template <int WARPS, int MINB, int DBG = 0>
__global__ void __launch_bounds__(WARPS * 32, MINB) nn_ring_kernel(const RingNNArgs a) {
    constexpr int RF_WARPS = WARPS, RF_ROWS = WARPS * RF_WROWS;  // a CTA owns WARPS x 256 rows (a.rowchunks is sized for it)
    __shared__ __align__(128) float rows_s[RF_ROWS * 3];
    __shared__ __align__(128) float cols_raw[RF_COLS * 3];
    __shared__ __align__(128) float cols_p[RF_COLS * 3];
    __shared__ __align__(16) float wmn_s[RF_WARPS][RF_COLS];  // per warp: running column minima ...
    __shared__ __align__(16) int wrot_s[RF_WARPS][RF_COLS];   // ... and the rotation that last improved them
    __shared__ __align__(16) u64 wcolkey[RF_WARPS][RF_COLS];  // per warp: best (distance, row) key of every column of the round
    __shared__ __align__(8) uint64_t mbar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    asm volatile("griddepcontrol.launch_dependents;");  // a programmatically dependent tail kernel may start its prologue
    HP_TRACE((size_t)blockIdx.x * 2);
    int bid = blockIdx.x;
    const int cc = bid % a.colchunks;
    bid /= a.colchunks;
    const int rc = bid % a.rowchunks;
    const int cloud = bid / a.rowchunks;
    const int n = a.n, m = a.m;
    const float *__restrict__ A = a.set1 + (size_t)cloud * n * 3;
    const float *__restrict__ Bp = a.set2 + (size_t)cloud * m * 3;
    const int row0 = rc * RF_ROWS;                       // first row of this CTA
    const int nrows = min(RF_ROWS, n - row0);            // real rows of this CTA (>= 1)

    // rows and the first round's columns arrive together: one mbarrier phase, two bulk copies in flight
    const int cbase0 = cc * a.R * RF_COLS;               // < m by construction of colchunks
    const int ncols0 = min(RF_COLS, m - cbase0);
    if (tid == 0) mbar_init(&mbar, 1);
    __syncthreads();
    uint32_t phase = 0;
    {
        const float *rsrc = A + (size_t)row0 * 3, *csrc = Bp + (size_t)cbase0 * 3;
        const uint32_t rb = rf_bulk_bytes(rsrc, nrows), cb = rf_bulk_bytes(csrc, ncols0);
        if (tid == 0 && rb + cb) {
            fence_proxy_async();
            mbar_expect_tx(&mbar, rb + cb);
            if (rb) bulk_g2s(rows_s, rsrc, rb, &mbar);
            if (cb) bulk_g2s(cols_raw, csrc, cb, &mbar);
        }
        // block-uniform: nothing to do when the bulk copies cover whole, unpadded tiles (the common case)
        if (rb != (uint32_t)RF_ROWS * 12u) rf_stage_tail(rows_s, rsrc, nrows, RF_ROWS, RF_PAD_ROW, rb, tid, RF_WARPS * 32);
        if (cb != (uint32_t)RF_COLS * 12u) rf_stage_tail(cols_raw, csrc, ncols0, RF_COLS, RF_PAD_COL, cb, tid, RF_WARPS * 32);
        if (rb + cb) {
            mbar_wait(&mbar, phase);
            phase ^= 1;
        }
    }
    __syncthreads();

    const int wrow0 = warp * RF_WROWS;                   // first (CTA-local) row of this warp
    const bool warp_active = wrow0 < nrows;              // warp-uniform
    const int lrow0 = wrow0 + lane * RF_RQ;              // first (CTA-local) row of this lane
    float qx[RF_RQ], qy[RF_RQ], qz[RF_RQ];
    {
        const float4 *rp = reinterpret_cast<const float4 *>(rows_s + lrow0 * 3);  // 24 floats, 16-byte aligned
        float v[24];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const float4 t4 = rp[i];
            v[4 * i + 0] = t4.x, v[4 * i + 1] = t4.y, v[4 * i + 2] = t4.z, v[4 * i + 3] = t4.w;
        }
#pragma unroll
        for (int j = 0; j < RF_RQ; ++j) qx[j] = v[3 * j], qy[j] = v[3 * j + 1], qz[j] = v[3 * j + 2];
    }
    float *wmn = wmn_s[warp];
    int *wrot = wrot_s[warp];

    for (int r = 0; r < a.R; ++r) {
        const int cbase = (cc * a.R + r) * RF_COLS;      // first column of this round
        if (cbase >= m) break;                           // uniform
        const int ncols = min(RF_COLS, m - cbase);
        // here: cols_raw holds this round's columns (all threads have waited + synchronised), previous round's state consumed
        {   // permute the columns into the packed group layout; reset the merge keys and this warp's column state
#pragma unroll
            for (int c = tid; c < RF_COLS; c += RF_WARPS * 32) {
                const int g = c >> 2, k = c & 3;
                float *dst = cols_p + g * 12 + ((k & 2) ? 6 : 0) + (k & 1);
                dst[0] = cols_raw[c * 3 + 0], dst[2] = cols_raw[c * 3 + 1], dst[4] = cols_raw[c * 3 + 2];
#pragma unroll
                for (int w = 0; w < RF_WARPS; ++w) wcolkey[w][c] = ~0ull;  // inactive warps stay at 'none'
            }
#pragma unroll
            for (int i = 0; i < RF_RC; ++i) wmn[lane * RF_RC + i] = __int_as_float(0x7f800000), wrot[lane * RF_RC + i] = 0;
        }
        __syncthreads();
        // cols_raw is free again: the next round's columns fly in under this round's rotations
        const int cbase_n = cbase + RF_COLS;
        const bool more = (r + 1 < a.R) && (cbase_n < m);  // uniform
        uint32_t nb_bytes = 0;
        if (more) {
            const int ncols_n = min(RF_COLS, m - cbase_n);
            const float *csrc = Bp + (size_t)cbase_n * 3;
            nb_bytes = rf_bulk_bytes(csrc, ncols_n);
            if (tid == 0 && nb_bytes) {
                fence_proxy_async();
                mbar_expect_tx(&mbar, nb_bytes);
                bulk_g2s(cols_raw, csrc, nb_bytes, &mbar);
            }
            if (nb_bytes != (uint32_t)RF_COLS * 12u) rf_stage_tail(cols_raw, csrc, ncols_n, RF_COLS, RF_PAD_COL, nb_bytes, tid, RF_WARPS * 32);
        }
        if (warp_active) {
            // lane L meets column group (31 - L + t) mod 32 at rotation t: ascending with one wrap (at t = L+1);
            // a group's running minimum is handed from lane to lane through shared memory: it visits lanes
            // (31 - g + t) mod 32, i.e. ascending rows, wrapping to the lowest rows when lane 0 picks it up.
            float best[RF_RQ];
            int rot[RF_RQ];
#pragma unroll
            for (int j = 0; j < RF_RQ; ++j) best[j] = __int_as_float(0x7f800000), rot[j] = 0;
            int g = 31 - lane;
            RFGroup cur = rf_load_group(cols_p, g);
#pragma unroll 2
            for (int t = 0; t < ((DBG & 1) ? 2 : 32); ++t) {
                const int gn = (g + 1) & 31;
                const RFGroup nxt = rf_load_group(cols_p, gn);  // prefetch the next rotation's columns
                float4 mnv = *reinterpret_cast<const float4 *>(wmn + g * RF_RC);
                int4 rtv = *reinterpret_cast<const int4 *>(wrot + g * RF_RC);
                {   // at t == lane+1 this lane's column groups wrap to index 0: later groups must win ties (+1 ulp)
                    const int wrapped = (t == lane + 1) ? 1 : 0;
#pragma unroll
                    for (int j = 0; j < RF_RQ; ++j) best[j] = __int_as_float(__float_as_int(best[j]) + wrapped);
                }
                {   // lane 0 (t >= 1) picks up columns last seen by lane 31: they wrap to the lowest rows (+1 ulp)
                    const int wrapped = (lane == 0 && t != 0) ? 1 : 0;
                    mnv.x = __int_as_float(__float_as_int(mnv.x) + wrapped), mnv.y = __int_as_float(__float_as_int(mnv.y) + wrapped);
                    mnv.z = __int_as_float(__float_as_int(mnv.z) + wrapped), mnv.w = __int_as_float(__float_as_int(mnv.w) + wrapped);
                }
                float cm[RF_RC] = {mnv.x, mnv.y, mnv.z, mnv.w};  // running column minima, continued from the handed-over state
#pragma unroll
                for (int j = 0; j < RF_RQ; j += 2) {
                    float d[2][4];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const f32x2 px = pack2(qx[j + u], qx[j + u]), py = pack2(qy[j + u], qy[j + u]), pz = pack2(qz[j + u], qz[j + u]);
                        unpack2(sqdist_exact2(px, py, pz, cur.x01, cur.y01, cur.z01), d[u][0], d[u][1]);
                        unpack2(sqdist_exact2(px, py, pz, cur.x23, cur.y23, cur.z23), d[u][2], d[u][3]);
                        const float old = best[j + u];
                        float nb = min3(old, d[u][0], d[u][1]);
                        nb = min3(nb, d[u][2], d[u][3]);
                        best[j + u] = nb;
                        rot[j + u] = (nb < old) ? t : rot[j + u];
                    }
#pragma unroll
                    for (int i = 0; i < RF_RC; ++i) cm[i] = min3(cm[i], d[0][i], d[1][i]);
                }
                rtv.x = (cm[0] < mnv.x) ? t : rtv.x, rtv.y = (cm[1] < mnv.y) ? t : rtv.y;
                rtv.z = (cm[2] < mnv.z) ? t : rtv.z, rtv.w = (cm[3] < mnv.w) ? t : rtv.w;
                mnv = make_float4(cm[0], cm[1], cm[2], cm[3]);
                *reinterpret_cast<float4 *>(wmn + g * RF_RC) = mnv;
                *reinterpret_cast<int4 *>(wrot + g * RF_RC) = rtv;
                __syncwarp();
                cur = nxt;
                g = gn;
            }


            if (DBG & 2) {  // timing experiment only: no re-evaluation, one dummy key per lane
                float sm = 0.f; int sr = 0;
#pragma unroll
                for (int j = 0; j < RF_RQ; ++j) sm += best[j], sr += rot[j];
                red_max_u64(a.rowkey + (size_t)cloud * n + row0 + lrow0, ((u64)__float_as_uint(sm) << 32) | (unsigned)sr);
            } else {
            // ---- rows: exact value, then the lowest column index of the winning group with d == value ----
            // (branch-free: the 8 re-evaluations are independent so that their loads and FMA chains overlap)
            u64 rkey[RF_RQ];
#pragma unroll
            for (int j = 0; j < RF_RQ; ++j) {
                float bv = best[j];
                if (lane <= 30 && rot[j] <= lane) bv = bump_down(bv);  // bumped at t = lane+1 and never replaced
                const int gw = (31 - lane + rot[j]) & 31;
                const float4 *cp = reinterpret_cast<const float4 *>(cols_p + gw * 12);
                const float4 v0 = cp[0], v1 = cp[1], v2 = cp[2];  // x0 x1 y0 y1 | z0 z1 x2 x3 | y2 y3 z2 z3
                const float d0 = sqdist_exact(qx[j], qy[j], qz[j], v0.x, v0.z, v1.x);
                const float d1 = sqdist_exact(qx[j], qy[j], qz[j], v0.y, v0.w, v1.y);
                const float d2 = sqdist_exact(qx[j], qy[j], qz[j], v1.z, v2.x, v2.z);
                const int k = (d0 == bv) ? 0 : (d1 == bv) ? 1 : (d2 == bv) ? 2 : 3;
                rkey[j] = ~(((u64)__float_as_uint(bv) << 32) | (unsigned)(cbase + gw * RF_RC + k));
            }
            {
                const int nvalid = n - (row0 + lrow0);            // real rows of this lane (may be <= 0 or >= 8)
                u64 *rk = a.rowkey + (size_t)cloud * n + row0 + lrow0;
#pragma unroll
                for (int j = 0; j < RF_RQ; ++j)
                    if (j < nvalid) red_max_u64(rk + j, rkey[j]);
            }
            // ---- columns: lane L finalises group L (padding columns included: their keys are never merged) ----
            {
                const float4 *cp = reinterpret_cast<const float4 *>(cols_p + lane * 12);
                const float4 v0 = cp[0], v1 = cp[1], v2 = cp[2];
                const float cxs[RF_RC] = {v0.x, v0.y, v1.z, v1.w}, cys[RF_RC] = {v0.z, v0.w, v2.x, v2.y},
                            czs[RF_RC] = {v1.x, v1.y, v2.z, v2.w};
                const float4 mnf = *reinterpret_cast<const float4 *>(wmn + lane * RF_RC);
                const int4 rtf = *reinterpret_cast<const int4 *>(wrot + lane * RF_RC);
                const float mns[RF_RC] = {mnf.x, mnf.y, mnf.z, mnf.w};
                const int rts[RF_RC] = {rtf.x, rtf.y, rtf.z, rtf.w};
                u64 ckey[RF_RC];
#pragma unroll
                for (int i = 0; i < RF_RC; ++i) {
                    float cv = mns[i];
                    if (lane <= 30 && rts[i] <= lane) cv = bump_down(cv);  // bumped when lane 0 picked it up, never replaced
                    const int vl = (31 - lane + rts[i]) & 31;              // lane whose rows produced the minimum
                    const float4 *rp = reinterpret_cast<const float4 *>(rows_s + (wrow0 + vl * RF_RQ) * 3);  // 8 rows, 96 B
                    float rv[24];
#pragma unroll
                    for (int q4 = 0; q4 < 6; ++q4) {
                        const float4 t4 = rp[q4];
                        rv[4 * q4 + 0] = t4.x, rv[4 * q4 + 1] = t4.y, rv[4 * q4 + 2] = t4.z, rv[4 * q4 + 3] = t4.w;
                    }
                    int k = RF_RQ - 1;
#pragma unroll
                    for (int q = RF_RQ - 2; q >= 0; --q) {
                        const float dq = sqdist_exact(rv[3 * q], rv[3 * q + 1], rv[3 * q + 2], cxs[i], cys[i], czs[i]);
                        k = (dq == cv) ? q : k;
                    }
                    ckey[i] = ((u64)__float_as_uint(cv) << 32) | (unsigned)(row0 + wrow0 + vl * RF_RQ + k);
                }
                ulonglong2 *ck = reinterpret_cast<ulonglong2 *>(&wcolkey[warp][lane * RF_RC]);
                ck[0] = make_ulonglong2(ckey[0], ckey[1]);
                ck[1] = make_ulonglong2(ckey[2], ckey[3]);
            }
            }
        }  // warp_active
        __syncthreads();
#pragma unroll
        for (int c = tid; c < RF_COLS; c += RF_WARPS * 32) {
            if (c < ncols) {
                u64 key = wcolkey[0][c];
#pragma unroll
                for (int w = 1; w < RF_WARPS; ++w) key = wcolkey[w][c] < key ? wcolkey[w][c] : key;
                red_max_u64(a.colkey + (size_t)cloud * m + cbase + c, ~key);
            }
        }
        if (more) {
            if (nb_bytes) {
                mbar_wait(&mbar, phase);
                phase ^= 1;
            }
            __syncthreads();                             // next columns visible, wcolkey / cols_p consumed
        }
    }
    // this CTA's candidates are merged: one more arrival on the cloud's ticket (bar.sync orders every thread's key atomics
    // before thread 0's fence; the tail kernel's CTAs of this cloud acquire the ticket before they read the keys)
    if (!(DBG & 4) && a.tickets) {  // (DBG 4: timing experiment without the arrival)
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            red_add_u32(a.ticket + cloud, 1u);
        }
    }
    HP_TRACE((size_t)blockIdx.x * 2 + 1);
}


// Stable inverse of an index map, for the atomic-free backward.  One CTA per (cloud, direction), at the tail of the
// unpack kernel.  `keys[e]` = target index of source e.  Output per cloud: perm[cnt] = sources sorted by (target, source),
// begin[ntgt] / end[ntgt] = every target's bucket in perm.
//   clouds up to 8192 points : stable radix sort of the composites (target << shift | source), see rf_build_inverse;
//   larger clouds            : bitonic sort of the composites.
// Both give the same, fully deterministic result whatever the bucket sizes are.

__device__ __forceinline__ void rf_inverse_bitonic(const RingNNArgs &a, unsigned int *sortbuf, const int *keys_or_null, int cnt,
                                                   int ntgt, int *perm, int *begin, int *end, int tid) {
    const int N = a.sort_n;
    if (keys_or_null != nullptr) {  // composites not yet built (coming from the fast path's layout)
        unsigned int tmp[RF_MAX_PER_THREAD];
#pragma unroll
        for (int q = 0; q < RF_MAX_PER_THREAD; ++q) {
            const int e = tid + q * RF_MERGE_THREADS;
            tmp[q] = (e < cnt) ? (((unsigned)keys_or_null[e] << a.sort_shift) | (unsigned)e) : 0xffffffffu;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < RF_MAX_PER_THREAD; ++q) {
            const int e = tid + q * RF_MERGE_THREADS;
            if (e < N) sortbuf[e] = tmp[q];
        }
    } else {
        for (int e = cnt + tid; e < N; e += RF_MERGE_THREADS) sortbuf[e] = 0xffffffffu;
    }
    for (int i = tid; i < 2 * ntgt; i += RF_MERGE_THREADS) begin[i] = 0;  // empty buckets: begin = end = 0 (end follows begin)
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < N; i += RF_MERGE_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned x = sortbuf[i], y = sortbuf[ixj];
                    const bool asc = (i & k) == 0;
                    if ((x > y) == asc) sortbuf[i] = y, sortbuf[ixj] = x;
                }
            }
            __syncthreads();
        }
    }
    const unsigned mask = (1u << a.sort_shift) - 1u;
    for (int p = tid; p < cnt; p += RF_MERGE_THREADS) {
        const unsigned c = sortbuf[p];
        const int key = (int)(c >> a.sort_shift);
        perm[p] = (int)(c & mask);
        if (p == 0 || (int)(sortbuf[p - 1] >> a.sort_shift) != key) begin[key] = p;
        if (p == cnt - 1 || (int)(sortbuf[p + 1] >> a.sort_shift) != key) end[key] = p + 1;
    }
}

// Fast path: stable LSD radix sort of the composites (target << shift | source) on the TARGET bits, 6 bits per pass.
// Stability keeps the sources of a bucket in ascending order, so no per-bucket sort is needed and the cost does not
// depend on how skewed the assignment is (early in training most points map to a handful of targets).
//   per pass: digit histogram (warp-aggregated) -> exclusive scan -> rounds of 1024 elements in order: every warp ranks
//   its elements among equal digits with __match_any_sync, a 64 x 32 (digit x warp) table turns the ranks into positions.
// smem (uints): buf0[cnt] | buf1[cnt] | wcnt[32][64] | dbase[64]
constexpr int RX_BITS = 6, RX_BINS = 1 << RX_BITS, RX_WARPS = RF_MERGE_THREADS / 32;

// `cur` holds the cnt composites, `alt` is the second buffer (both `big` uints), `wcnt` the rank table + digit starts.
__device__ __forceinline__ void rf_inverse_radix(const RingNNArgs &a, unsigned int *cur, unsigned int *alt, unsigned int *wcnt, int cnt,
                                                 int ntgt, int *perm, int *begin, int *end, int tid) {
    unsigned int *dbase = wcnt + RX_WARPS * RX_BINS;  // [RX_BINS] running start of every digit
    const int lane = tid & 31, warp = tid >> 5;
    const int shift0 = a.sort_shift;                  // composites: target << shift0 | source
    int tbits = 0;
    while ((1 << tbits) < ntgt) ++tbits;
    for (int i = tid; i < 2 * ntgt; i += RF_MERGE_THREADS) begin[i] = 0;  // empty buckets: begin = end = 0
    const int rounds = (cnt + RF_MERGE_THREADS - 1) / RF_MERGE_THREADS;
    for (int sh = shift0; sh < shift0 + tbits; sh += RX_BITS) {
        if (tid < RX_BINS) dbase[tid] = 0u;
        __syncthreads();  // composites / previous pass visible
        for (int r = 0; r < rounds; ++r) {  // histogram, one aggregated atomic per (warp, digit)
            const int e = r * RF_MERGE_THREADS + tid;
            const int d = e < cnt ? (int)((cur[e] >> sh) & (RX_BINS - 1)) : RX_BINS;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            if (d < RX_BINS && lane == __ffs(peers) - 1) atomicAdd(&dbase[d], (unsigned)__popc(peers));
        }
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the 64 digit counts (two per lane)
            const unsigned c0 = dbase[2 * lane], c1 = dbase[2 * lane + 1];
            unsigned inc = c0 + c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            const unsigned ex = inc - (c0 + c1);
            dbase[2 * lane] = ex, dbase[2 * lane + 1] = ex + c0;
        }
        for (int r = 0; r < rounds; ++r) {
            for (int i = tid; i < RX_WARPS * RX_BINS; i += RF_MERGE_THREADS) wcnt[i] = 0u;
            __syncthreads();  // table cleared; (first round) scan visible; previous round's scatter done
            const int e = r * RF_MERGE_THREADS + tid;
            const unsigned c = e < cnt ? cur[e] : 0u;
            const int d = e < cnt ? (int)((c >> sh) & (RX_BINS - 1)) : RX_BINS;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            if (d < RX_BINS && rank == 0) wcnt[warp * RX_BINS + d] = (unsigned)__popc(peers);
            __syncthreads();
            // per digit: exclusive scan over the 32 warps (one warp scans two digits, lane = source warp) -> positions
#pragma unroll
            for (int q = 0; q < RX_BINS / RX_WARPS; ++q) {
                const int dg = warp * (RX_BINS / RX_WARPS) + q;
                const unsigned v = wcnt[lane * RX_BINS + dg];
                unsigned inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += u;
                }
                const unsigned run = dbase[dg];
                wcnt[lane * RX_BINS + dg] = run + inc - v;
                __syncwarp();
                if (lane == 31) dbase[dg] = run + inc;
            }
            __syncthreads();
            if (d < RX_BINS) alt[wcnt[warp * RX_BINS + d] + rank] = c;
            __syncthreads();  // positions consumed before the table is cleared again
        }
        unsigned int *t = cur;
        cur = alt, alt = t;
    }
    __syncthreads();
    const unsigned mask = (1u << shift0) - 1u;
    for (int p = tid; p < cnt; p += RF_MERGE_THREADS) {
        const unsigned c = cur[p];
        const int key = (int)(c >> shift0);
        perm[p] = (int)(c & mask);
        if (p == 0 || (int)(cur[p - 1] >> shift0) != key) begin[key] = p;
        if (p == cnt - 1 || (int)(cur[p + 1] >> shift0) != key) end[key] = p + 1;
    }
}

// Entry: counting sort when the assignment is benign (every bucket <= RF_BUCKET_MAX sources: histogram, scan, placement
// in arrival order, owner sorts its tiny bucket), stable radix sort otherwise.  smem (ints):
//   keys[cnt] | count[ntgt] | cursor[ntgt] | pbuf[cnt]   (<= 4*big)  followed by the radix rank table (RX_WARPS*RX_BINS + RX_BINS);
//   the radix fallback re-uses count.. as its two composite buffers (2*big <= 3*big).
constexpr int RF_BUCKET_MAX = 48;

// `perm` [cnt], `begin` [ntgt] and `end` [ntgt] (begin and end contiguous) may live in global OR shared memory.
__device__ __forceinline__ void rf_build_inverse(const RingNNArgs &a, unsigned int *smem_u, bool dir2, int tid, int *perm,
                                                 int *begin, int *end) {
    __shared__ int scan_warp[RF_MERGE_THREADS / 32];
    __shared__ int max_bucket;
    const int cnt = dir2 ? a.m : a.n;      // sources: the points whose nearest neighbour was searched
    const int ntgt = dir2 ? a.n : a.m;     // targets: the points they can map to
    if (!a.inv_fast) {  // big clouds: bitonic sort of the composites the caller wrote into the buffer
        rf_inverse_bitonic(a, smem_u, nullptr, cnt, ntgt, perm, begin, end, tid);
        return;
    }
    const int big = a.n > a.m ? a.n : a.m;
    int *keys = reinterpret_cast<int *>(smem_u);
    int *count = keys + big, *cursor = count + big, *pbuf = cursor + big;
    unsigned int *table = smem_u + 4 * (size_t)big;
    for (int i = tid; i < ntgt; i += RF_MERGE_THREADS) count[i] = 0;
    if (tid == 0) max_bucket = 0;
    __syncthreads();
    for (int e = tid; e < cnt; e += RF_MERGE_THREADS) atomicAdd(&count[keys[e]], 1);
    __syncthreads();
    // exclusive scan of count[0..ntgt): contiguous chunk per thread, warp scan, scan of the warp totals
    const int per = (ntgt + RF_MERGE_THREADS - 1) / RF_MERGE_THREADS;
    const int lo = min(ntgt, tid * per), hi = min(ntgt, lo + per);
    int local = 0, lmax = 0;
    for (int i = lo; i < hi; ++i) local += count[i], lmax = max(lmax, count[i]);
    int inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if ((tid & 31) >= o) inc += v;
    }
    if ((tid & 31) == 31) scan_warp[tid >> 5] = inc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    if ((tid & 31) == 0 && lmax > 0) atomicMax(&max_bucket, lmax);
    __syncthreads();
    if (tid < 32) {
        int w = scan_warp[tid], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, winc, o);
            if (tid >= o) winc += v;
        }
        scan_warp[tid] = winc - w;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    if (max_bucket > RF_BUCKET_MAX) {  // block-uniform: skewed assignment -> data-independent radix sort (keys are intact)
        unsigned int *buf0 = reinterpret_cast<unsigned int *>(count), *buf1 = buf0 + big;
        __syncthreads();
        for (int e = tid; e < cnt; e += RF_MERGE_THREADS) buf0[e] = ((unsigned)keys[e] << a.sort_shift) | (unsigned)e;
        rf_inverse_radix(a, buf0, buf1, table, cnt, ntgt, perm, begin, end, tid);
        return;
    }
    {
        int run = scan_warp[tid >> 5] + inc - local;
        for (int i = lo; i < hi; ++i) {
            const int c = count[i];
            count[i] = run;   // begin
            cursor[i] = run;
            begin[i] = c ? run : 0;  // empty buckets: begin = end = 0
            end[i] = c ? run + c : 0;
            run += c;
        }
    }
    __syncthreads();
    for (int e = tid; e < cnt; e += RF_MERGE_THREADS) pbuf[atomicAdd(&cursor[keys[e]], 1)] = e;  // arrival order
    __syncthreads();
    for (int i = tid; i < ntgt; i += RF_MERGE_THREADS) {  // owner sorts its bucket by source index (insertion sort)
        const int b0 = count[i], b1 = cursor[i];
        for (int p = b0 + 1; p < b1; ++p) {
            const int v = pbuf[p];
            int q = p - 1;
            while (q >= b0 && pbuf[q] > v) pbuf[q + 1] = pbuf[q], --q;
            pbuf[q + 1] = v;
        }
    }
    __syncthreads();
    for (int p = tid; p < cnt; p += RF_MERGE_THREADS) perm[p] = pbuf[p];
}


// ---- deterministic loss ------------------------------------------------------------------------------------------
// One summation order for every path (bit-identical losses from the unpack kernel and from the tail kernel):
//   chunk partial            = the 256 distances of sources [256r, 256r + 256): xor-shuffle tree per warp, then the 8 warp
//                              sums serially in warp order;
//   (cloud, direction) sum   = its chunk partials serially in ascending order;
//   total                    = thread t of 256 adds the (cloud, direction) sums t, t + 256, ... serially; the 256 accumulators
//                              are folded by the xor tree per warp and serially over the 8 warps.
// The partials meet in a[...].losspart; the last CTA to arrive on counters[0] folds them and restores the zero state.
constexpr int LS_CHUNK = 256;

__device__ __forceinline__ float rf_serial8(const float *w) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += w[i];
    return s;
}

// Called by ONE WARP of the last-arriving CTA; lane l plays threads l, l + 32, ..., l + 224 of the 256 of the definition above.
// chunks1 / chunks2 > 0: losspart holds chunk partials, per cloud [chunks1 of direction 0 | chunks2 of direction 1];
// chunks1 == 0: losspart holds one sum per (cloud, direction).
__device__ __forceinline__ void rf_loss_total_warp(const RingNNArgs &a, int chunks1, int chunks2, int lane) {
    float accv[8];
#pragma unroll
    for (int vw = 0; vw < 8; ++vw) {  // "warp" vw of the definition; unrolled: the loads of the eight groups overlap
        float acc = 0.f;
        for (int pd = vw * 32 + lane; pd < 2 * a.b; pd += 256) {
            float p;
            if (chunks1 == 0) {
                p = __ldcg(a.losspart + pd);
                a.losspart[pd] = 0.f;
            } else {
                float *q = a.losspart + (size_t)(pd >> 1) * (chunks1 + chunks2) + ((pd & 1) ? chunks1 : 0);
                const int nch = (pd & 1) ? chunks2 : chunks1;
                p = 0.f;
                for (int r0 = 0; r0 < nch; r0 += 8) {  // eight loads in flight, then the serial sum in chunk order
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = (r0 + i < nch) ? __ldcg(q + r0 + i) : 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (r0 + i < nch) {
                            p += v[i];
                            q[r0 + i] = 0.f;
                        }
                    }
                }
            }
            acc += p;
        }
        accv[vw] = acc;
    }
    float tot = 0.f;
#pragma unroll
    for (int vw = 0; vw < 8; ++vw) tot += warp_sum(accv[vw]);  // serial over the 8 "warps", each folded by the xor tree
    if (lane == 0) {
        a.loss[0] = tot;
        a.counters[0] = 0u;
    }
}

// key -> (distance, index); restores the zero state of the key arrays and of the cloud's ring ticket; fixed-order loss.
// One block per (cloud, direction).
template <bool INVERT>
__global__ void __launch_bounds__(RF_MERGE_THREADS) nn_ring_unpack_kernel(const RingNNArgs a) {
    extern __shared__ __align__(16) unsigned int sortbuf[];  // INVERT: [sort_n] composites
    __shared__ float warp_part[RF_MERGE_THREADS / 32];
    __shared__ int flag;
    const int tid = threadIdx.x;
    const int cloud = blockIdx.x >> 1;
    const bool dir2 = blockIdx.x & 1;
    const int cnt = dir2 ? a.m : a.n;
    u64 *src = (dir2 ? a.colkey : a.rowkey) + (size_t)cloud * cnt;
    float *dist = (dir2 ? a.dist2 : a.dist1) + (size_t)cloud * cnt;
    int *idx = (dir2 ? a.idx2 : a.idx1) + (size_t)cloud * cnt;
    // launched programmatically dependent on the ring kernel: resident while that drains, running once its keys are complete
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!dir2 && tid == 0 && a.tickets) a.ticket[cloud] = 0u;  // the ring kernel's arrivals are not needed on this path
    // Phase 1 only READS the keys: the device-scope fence of the loss ticket below then has no stores of this SM to
    // drain (a fence issued after the 3 stores per element costs ~10 us).  Phase 2 (unpack_store) writes the outputs.
    auto unpack_store = [&]() {
        for (int e = tid; e < cnt; e += RF_MERGE_THREADS) {
            const u64 key = ~__ldcg(src + e);
            src[e] = 0ull;
            dist[e] = __uint_as_float((unsigned)(key >> 32));
            idx[e] = (int)(unsigned)(key & 0xffffffffu);
            if (INVERT) sortbuf[e] = a.inv_fast ? (unsigned)(key & 0xffffffffu) : (((unsigned)(key & 0xffffffffu) << a.sort_shift) | (unsigned)e);
        }
        if (INVERT) {
            const int ntgt = dir2 ? a.n : a.m;
            int *inv = (dir2 ? a.inv2 : a.inv1) + (size_t)cloud * (cnt + 2 * ntgt);
            rf_build_inverse(a, sortbuf, dir2, tid, inv, inv + cnt, inv + cnt + ntgt);
        }
    };
    if (a.loss == nullptr) {
        unpack_store();
        return;
    }
    float p = 0.f;  // thread 0: this (cloud, direction)'s sum, chunk by chunk (see "deterministic loss")
    for (int base = 0; base < cnt; base += RF_MERGE_THREADS) {  // block-uniform trip count
        const int e = base + tid;
        float v = e < cnt ? __uint_as_float((unsigned)((~__ldcg(src + e)) >> 32)) : 0.f;
        v = warp_sum(v);
        __syncthreads();  // warp_part reusable
        if ((tid & 31) == 0) warp_part[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int grp = 0; grp < RF_MERGE_THREADS / LS_CHUNK; ++grp)
                if (base + grp * LS_CHUNK < cnt) p += rf_serial8(warp_part + grp * 8);
        }
    }
    if (tid == 0) {
        a.losspart[blockIdx.x] = p;
        __threadfence();
        flag = (atomicAdd(a.counters, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (flag && tid < 32) {  // last arriver: one warp folds the (cloud, direction) sums
        __threadfence();
        rf_loss_total_warp(a, 0, 0, tid);
    }
    unpack_store();
}

// ---- fused tail of the training step: unpack + loss + inverse index maps + BOTH gradients ------------------------------
// One CTA of 256 threads per (cloud, direction, section of 256 TARGET points); thread = target.
//     direction 0: rows -> columns (idx1), targets = columns,  grad_xyz2[k] = 2g[(b_k - a_idx2[k]) + sum_{i: idx1[i]=k} (b_k - a_i)]
//     direction 1: columns -> rows (idx2), targets = rows,     grad_xyz1[i] = 2g[(a_i - b_idx1[i]) + sum_{k: idx2[k]=i} (a_i - b_k)]
// (nndistance.cu:143-151).  The kernel is launched programmatically dependent on the ring kernel and does NOT wait for the
// whole grid: the ring kernel's CTAs are ordered cloud-major and bump a per-cloud ticket when their keys are merged; a tail
// CTA stages its points and clears its tables while it waits for ITS cloud's ticket, so the tails of the early clouds run
// under the ring kernel's second wave and only the last clouds' tails (a few microseconds, all in parallel) are exposed.
//   * every CTA reads all `cnt` source keys of its direction (8 per thread at 2048 points) and keeps those whose target falls
//     into its section.  The inverse map is a STABLE counting sort without any sorting: sources are ranked among the equal
//     targets of their 32-source chunk by __match_any_sync, a [chunk][target] count table is prefix-summed per target over
//     the chunks (thread = target), and position = bucket begin + chunk prefix + rank.  Data-independent cost, buckets in
//     ascending source order whatever the skew of the assignment.
//   * gradients: arithmetic and summation order of nn_grad_gather_kernel (ascending source index; buckets above GATHER_COOP
//     warp-cooperatively), so they are bit-identical to the three-kernel path.
//   * distances / indices / loss chunk r of the direction are written by section r mod (number of sections).
//   * the last tail CTA of a cloud to have read the keys (done ticket) restores the zero state of both key arrays and of the
//     cloud's tickets; the last CTA on the loss ticket folds the loss and waits for the ring grid (griddepcontrol.wait), so
//     the tail grid never completes before its prerequisite.
struct RingTailArgs {
    const float *g;          // device scalar: upstream gradient of the loss
    float *grad1, *grad2;    // [b,n,3], [b,m,3]
    int sec1, sec2;          // sections of direction 0 (ceil(m / 256) column sections) and 1 (ceil(n / 256) row sections)
    int chunks1, chunks2;    // 256-source loss chunks of direction 0 (ceil(n / 256)) and 1 (ceil(m / 256))
    unsigned int expected;   // ring CTAs per cloud
};
constexpr int TL_THREADS = 256, TL_SEC = 256, TL_WARPS = TL_THREADS / 32;  // compute threads: thread = target of the section
constexpr int TL_CTA_THREADS = TL_THREADS + 32;                            // + one service warp
constexpr int TL_MAX_POINTS = 4096;  // keys of a direction live in registers: <= 16 per thread
constexpr int GATHER_COOP_F = 32;    // == GATHER_COOP of nn_grad_gather_kernel (same summation tree)
constexpr int TL_FAST = 8;           // buckets up to this size take the slot path (no ranking, no prefix sums)

// dynamic shared memory of the tail kernel; 0 if the shape is not supported (caller takes the three-kernel path)
static size_t rf_tail_smem_bytes(int n, int m, bool stage_sources = true) {
    const size_t big = (size_t)(n > m ? n : m);
    if (big > TL_MAX_POINTS) return 0;
    const size_t c32 = (big + 31) / 32;
    size_t bytes = c32 * TL_SEC * 2;               // table  u16 [chunk of 32 sources][target of the section]
    bytes += ((big * 2 + 15) & ~(size_t)15) * 2;   // info, pbuf  u16 [source]
    bytes += TL_SEC * 12;                          // target points of the section
    if (stage_sources) bytes += (big * 12 + 15) & ~(size_t)15;  // source points (last: absent in whole-grid mode)
    return bytes;
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// named barriers of the tail kernel: 0 = whole CTA (compute warps + service warp), 1 = the 8 compute warps,
// 2 = hand-over from the compute warps (arrive) to the service warp (sync)
__device__ __forceinline__ void tl_compute_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void tl_handover_arrive() { asm volatile("bar.arrive 2, 288;" ::: "memory"); }
__device__ __forceinline__ void tl_handover_wait() { asm volatile("bar.sync 2, 288;" ::: "memory"); }

template <int ROUNDS>  // ceil(cnt / 256) <= ROUNDS: the keys of the direction stay in registers between the phases
__global__ void __maxnreg__(56) nn_ring_tail_kernel(const RingNNArgs a, const RingTailArgs f) {
    extern __shared__ __align__(16) unsigned char tsm[];
    __shared__ float wp_s[ROUNDS][TL_WARPS];
    __shared__ int wscan_s[TL_WARPS];
    __shared__ int begin_s[TL_SEC];
    __shared__ int idxT_s[TL_SEC];
    __shared__ int flag_s[2];
    __shared__ int fcnt_s[TL_SEC];                                   // fast path: sources per target ...
    __shared__ __align__(16) unsigned short fslot_s[TL_SEC][TL_FAST];  // ... and the first TL_FAST of them, in arrival order
    __shared__ int fover_s;                                          // some target of the section has more than TL_FAST sources
    __shared__ __align__(8) uint64_t mbar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool service = warp == TL_WARPS;  // warp 8: staging, tickets, loss fold -- everything that waits on global memory
    const int per_cloud = f.sec1 + f.sec2;
    // cloud-major like the ring kernel: the tails of the early clouds (complete long before) run first and free their slots,
    // the tails of the last clouds become resident while the ring kernel drains and wait for their tickets
    const int cloud = blockIdx.x / per_cloud, rr = blockIdx.x - cloud * per_cloud;
    const bool dir2 = rr >= f.sec1;
    const int sec = dir2 ? rr - f.sec1 : rr, nsec = dir2 ? f.sec2 : f.sec1;
    const int cnt = dir2 ? a.m : a.n;       // sources of this direction
    const int ntgt = dir2 ? a.n : a.m;      // targets (the side whose gradient this CTA produces)
    const int tbase = sec * TL_SEC, tcount = min(TL_SEC, ntgt - tbase);
    const int big = a.n > a.m ? a.n : a.m;
    const int c32 = (cnt + 31) >> 5;
    unsigned short *table = reinterpret_cast<unsigned short *>(tsm);
    size_t off = (size_t)((big + 31) >> 5) * TL_SEC * 2;
    unsigned short *info = reinterpret_cast<unsigned short *>(tsm + off);
    off += ((size_t)big * 2 + 15) & ~(size_t)15;
    unsigned short *pbuf = reinterpret_cast<unsigned short *>(tsm + off);
    off += ((size_t)big * 2 + 15) & ~(size_t)15;
    float *T_s = reinterpret_cast<float *>(tsm + off);
    off += TL_SEC * 12;
    const float *__restrict__ Tg = (dir2 ? a.set1 : a.set2) + ((size_t)cloud * ntgt + tbase) * 3;
    const float *__restrict__ Sg = (dir2 ? a.set2 : a.set1) + (size_t)cloud * cnt * 3;
    float *S_stage = reinterpret_cast<float *>(tsm + off);                 // only carved out when a.stage_sources
    const float *__restrict__ S_s = a.stage_sources ? S_stage : Sg;        // sources: shared memory, or global through L1
    u64 *src = (dir2 ? a.colkey : a.rowkey) + (size_t)cloud * cnt;
    u64 *oth = (dir2 ? a.rowkey : a.colkey) + (size_t)cloud * ntgt;

    // ---- prologue (independent of the ring kernel's results): points into shared memory, count table cleared ----
    const size_t tr0 = (size_t)2 * a.b * a.rowchunks * a.colchunks + (size_t)blockIdx.x * 10;  // bench build: 10 timeline slots per CTA
    (void)tr0;
    HP_TRACE(tr0);
    if (tid == 0) mbar_init(&mbar, 1);
    __syncthreads();
    const uint32_t tb = rf_bulk_bytes(Tg, tcount), sb = a.stage_sources ? rf_bulk_bytes(Sg, cnt) : 0u;
    if (service) {
        if (lane == 0 && tb + sb) {
            fence_proxy_async();
            mbar_expect_tx(&mbar, tb + sb);
            if (tb) bulk_g2s(T_s, Tg, tb, &mbar);
            if (sb) bulk_g2s(S_stage, Sg, sb, &mbar);
        }
        // ---- wait for this cloud's ring CTAs (all of them are resident or done when this grid is allowed to start) ----
        if (f.expected == 0) {
            asm volatile("griddepcontrol.wait;" ::: "memory");  // whole-grid mode: the ring kernel's keys are complete and visible
        } else if (lane == 0) {
            while (ld_acquire_u32(a.ticket + cloud) < f.expected) __nanosleep(20);
        }
        HP_TRACE_T(TL_THREADS, tr0 + 1);
        __syncwarp();
    } else {
        if (tb != (uint32_t)tcount * 12u) rf_stage_tail(T_s, Tg, tcount, tcount, 0.f, tb, tid, TL_THREADS);  // uniform; rare
        if (a.stage_sources && sb != (uint32_t)cnt * 12u) rf_stage_tail(S_stage, Sg, cnt, cnt, 0.f, sb, tid, TL_THREADS);
        uint4 *t4 = reinterpret_cast<uint4 *>(table);
        for (int i = tid; i < c32 * (TL_SEC * 2 / 16); i += TL_THREADS) t4[i] = make_uint4(0u, 0u, 0u, 0u);
        fcnt_s[tid] = 0;
        if (tid == 0) fover_s = 0;
    }
    __syncthreads();  // ticket acquired (service warp), table cleared (compute warps)

    if (service) {
        // ---- service warp: done ticket, loss partials + loss ticket, total fold -- off the compute warps' critical path ----
        tl_handover_wait();  // the compute warps have read every key they need and written their loss warp sums
        if (lane == 0) {
            // both arrivals are issued before either result is looked at: the round trip of the first hides under the fence
            const unsigned old_done = atomicAdd(a.done + cloud, 1u);  // last reader of the cloud restores the keys
            unsigned old_cnt = 0xffffffffu;
            if (a.loss != nullptr) {
                const int nch = dir2 ? f.chunks2 : f.chunks1;
                float *lp = a.losspart + (size_t)cloud * (f.chunks1 + f.chunks2) + (dir2 ? f.chunks1 : 0);
                for (int r = sec; r < nch; r += nsec) lp[r] = rf_serial8(wp_s[r]);
                __threadfence();
                old_cnt = atomicAdd(a.counters, 1u);
            }
            flag_s[0] = old_done == (unsigned)per_cloud - 1;
            flag_s[1] = old_cnt == gridDim.x - 1;
        }
        HP_TRACE_T(TL_THREADS, tr0 + 7);
        __syncwarp();
        if (flag_s[1]) {  // warp-uniform: every tail CTA's partials are visible (their fence + ticket, ours below)
            __threadfence();
            rf_loss_total_warp(a, f.chunks1, f.chunks2, lane);
        }
    } else {
        // ---- keys: own direction (distance + target of every source), other direction (the own-term index of my target) ----
        u64 key[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int e = r * TL_THREADS + tid;
            key[r] = e < cnt ? ~__ldcg(src + e) : 0ull;
        }
        const u64 okey = tid < tcount ? ~__ldcg(oth + tbase + tid) : 0ull;
        if (a.loss != nullptr) {
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) {
                if (r * TL_THREADS < cnt && (r % nsec) == sec) {  // block-uniform: this section owns loss chunk r
                    const int e = r * TL_THREADS + tid;
                    const float v = warp_sum(e < cnt ? __uint_as_float((unsigned)(key[r] >> 32)) : 0.f);
                    if (lane == 0) wp_s[r][warp] = v;
                }
            }
        }
        if (tid < tcount) idxT_s[tid] = (int)(unsigned)(okey & 0xffffffffu);
        __threadfence_block();
        HP_TRACE(tr0 + 2);
        tl_handover_arrive();  // every key this CTA needs is in registers: the service warp may take the tickets (non-blocking for us)
        const float gs = __ldg(f.g) * 2.f;
        float *G = (dir2 ? f.grad1 : f.grad2) + ((size_t)cloud * ntgt + tbase + tid) * 3;
        // ---- slot path: when no target of the section has more than TL_FAST sources (every benign assignment: trained networks,
        //      uniform clouds), the buckets are filled by shared-memory atomics in arrival order and the target's thread orders its
        //      (at most TL_FAST) entries itself -- same ascending summation order, hence the same bits, as the general path ----
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int e = r * TL_THREADS + tid;
            const unsigned tl = (unsigned)(key[r] & 0xffffffffu) - (unsigned)tbase;
            if (e < cnt && tl < (unsigned)tcount) {
                const int pos = atomicAdd(&fcnt_s[tl], 1);
                if (pos < TL_FAST) fslot_s[tl][pos] = (unsigned short)e;
                else fover_s = 1;
            }
        }
        if (tb + sb) mbar_wait(&mbar, 0);
        tl_compute_sync();  // slots complete, points visible
        if (!fover_s) {     // block-uniform
            if (tid < tcount) {
                const float px = T_s[3 * tid + 0], py = T_s[3 * tid + 1], pz = T_s[3 * tid + 2];
                const int j = min(max(idxT_s[tid], 0), cnt - 1);
                float ax = gs * (px - S_s[3 * j + 0]), ay = gs * (py - S_s[3 * j + 1]), az = gs * (pz - S_s[3 * j + 2]);
                const int sz = fcnt_s[tid];
                const uint4 sl = *reinterpret_cast<const uint4 *>(fslot_s[tid]);  // 8 x u16
                int es[TL_FAST] = {(int)(sl.x & 0xffffu), (int)(sl.x >> 16), (int)(sl.y & 0xffffu), (int)(sl.y >> 16),
                                   (int)(sl.z & 0xffffu), (int)(sl.z >> 16), (int)(sl.w & 0xffffu), (int)(sl.w >> 16)};
#pragma unroll
                for (int q = 0; q < TL_FAST; ++q) es[q] = q < sz ? es[q] : 0x7fffffff;
                // Batcher's odd-even merge sort for 8 keys (19 compare-exchanges): ascending source order
#define HP_CE(i, j) { const int lo_ = min(es[i], es[j]); es[j] = max(es[i], es[j]); es[i] = lo_; }
                HP_CE(0, 1) HP_CE(2, 3) HP_CE(4, 5) HP_CE(6, 7)
                HP_CE(0, 2) HP_CE(1, 3) HP_CE(4, 6) HP_CE(5, 7)
                HP_CE(1, 2) HP_CE(5, 6)
                HP_CE(0, 4) HP_CE(1, 5) HP_CE(2, 6) HP_CE(3, 7)
                HP_CE(2, 4) HP_CE(3, 5)
                HP_CE(1, 2) HP_CE(3, 4) HP_CE(5, 6)
#undef HP_CE
#pragma unroll
                for (int q = 0; q < TL_FAST; ++q) {
                    if (q < sz) {
                        const int e = es[q];
                        ax += -(gs * (S_s[3 * e + 0] - px));
                        ay += -(gs * (S_s[3 * e + 1] - py));
                        az += -(gs * (S_s[3 * e + 2] - pz));
                    }
                }
                G[0] = ax, G[1] = ay, G[2] = az;
            }
        } else {
        // ---- general path: stable counting sort by ranking + chunk prefix sums (data-independent cost, any skew) ----
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int e = r * TL_THREADS + tid;
            if (r * TL_THREADS < cnt) {  // block-uniform
                const unsigned tl = (unsigned)(key[r] & 0xffffffffu) - (unsigned)tbase;
                const bool insec = e < cnt && tl < (unsigned)tcount;
                const unsigned peers = __match_any_sync(0xffffffffu, insec ? tl : 0xffffffffu);
                const unsigned rank = __popc(peers & ((1u << lane) - 1u));
                if (insec && rank == 0) table[(size_t)(r * TL_WARPS + warp) * TL_SEC + tl] = (unsigned short)__popc(peers);
                if (e < cnt) info[e] = insec ? (unsigned short)(tl | (rank << 8)) : (unsigned short)0xffffu;
            }
        }
        tl_compute_sync();     // table, info, idxT complete
        HP_TRACE(tr0 + 3);

        // ---- per target: prefix of its counts over the chunks (-> stable position of every source), bucket size ----
        int size = 0;
        {
            unsigned short *col = table + tid;
#pragma unroll 8
            for (int c = 0; c < c32; ++c) {
                const int v = col[(size_t)c * TL_SEC];
                col[(size_t)c * TL_SEC] = (unsigned short)size;
                size += v;
            }
        }
        int inc = size;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) wscan_s[warp] = inc;
        tl_compute_sync();  // warp totals visible
        int begin = inc - size;
#pragma unroll
        for (int w = 0; w < TL_WARPS; ++w) begin += (w < warp) ? wscan_s[w] : 0;
        begin_s[tid] = begin;
        tl_compute_sync();  // begin_s, chunk prefixes complete
        HP_TRACE(tr0 + 4);
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int e = r * TL_THREADS + tid;
            if (e < cnt) {
                const unsigned inf = info[e];
                if (inf != 0xffffu) {
                    const unsigned tl = inf & 255u;
                    pbuf[begin_s[tl] + table[(size_t)(e >> 5) * TL_SEC + tl] + (inf >> 8)] = (unsigned short)e;
                }
            }
        }
        tl_compute_sync();  // buckets complete
        HP_TRACE(tr0 + 5);

        // ---- gradient of my target: own term + its bucket in ascending source order ----
        {
            const bool valid = tid < tcount;
            float px = 0.f, py = 0.f, pz = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
            if (valid) {
                px = T_s[3 * tid + 0], py = T_s[3 * tid + 1], pz = T_s[3 * tid + 2];
                const int j = min(max(idxT_s[tid], 0), cnt - 1);
                ax = gs * (px - S_s[3 * j + 0]);
                ay = gs * (py - S_s[3 * j + 1]);
                az = gs * (pz - S_s[3 * j + 2]);
            }
            const int pb = begin, pe = begin + size;
            const bool bigb = valid && size > GATHER_COOP_F;
            if (valid && !bigb) {
                for (int p = pb; p < pe; ++p) {  // ascending source index: fixed summation order
                    const int e = pbuf[p];
                    ax += -(gs * (S_s[3 * e + 0] - px));
                    ay += -(gs * (S_s[3 * e + 1] - py));
                    az += -(gs * (S_s[3 * e + 2] - pz));
                }
            }
            unsigned todo = __ballot_sync(0xffffffffu, bigb);
            while (todo) {  // big buckets: the whole warp, lane-strided ascending + fixed shuffle tree
                const int srcl = __ffs(todo) - 1;
                todo &= todo - 1;
                const int b0 = __shfl_sync(0xffffffffu, pb, srcl), b1 = __shfl_sync(0xffffffffu, pe, srcl);
                const float qx = __shfl_sync(0xffffffffu, px, srcl), qy = __shfl_sync(0xffffffffu, py, srcl), qz = __shfl_sync(0xffffffffu, pz, srcl);
                float sx = 0.f, sy = 0.f, sz = 0.f;
                for (int p = b0 + lane; p < b1; p += 32) {
                    const int e = pbuf[p];
                    sx += -(gs * (S_s[3 * e + 0] - qx));
                    sy += -(gs * (S_s[3 * e + 1] - qy));
                    sz += -(gs * (S_s[3 * e + 2] - qz));
                }
                sx = warp_sum(sx), sy = warp_sum(sy), sz = warp_sum(sz);
                if (lane == srcl) ax += sx, ay += sy, az += sz;
            }
            if (valid) G[0] = ax, G[1] = ay, G[2] = az;
        }
        }  // general path
        // ---- distances / indices of the source chunks this section owns ----
        {
            float *dist = (dir2 ? a.dist2 : a.dist1) + (size_t)cloud * cnt;
            int *idx = (dir2 ? a.idx2 : a.idx1) + (size_t)cloud * cnt;
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) {
                const int e = r * TL_THREADS + tid;
                if ((r % nsec) == sec && e < cnt) {
                    dist[e] = __uint_as_float((unsigned)(key[r] >> 32));
                    idx[e] = (int)(unsigned)(key[r] & 0xffffffffu);
                }
            }
        }
    }
    HP_TRACE(tr0 + 6);
    __syncthreads();  // flags of the service warp visible to everybody
    if (flag_s[0]) {  // block-uniform: every tail CTA of the cloud has read the keys -> zero state for the next launch
        if (((cnt | ntgt) & 1) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(oth)) & 15) == 0) {
            ulonglong2 *z1 = reinterpret_cast<ulonglong2 *>(src), *z2 = reinterpret_cast<ulonglong2 *>(oth);
            for (int i = tid; i < cnt / 2; i += TL_CTA_THREADS) z1[i] = make_ulonglong2(0ull, 0ull);
            for (int i = tid; i < ntgt / 2; i += TL_CTA_THREADS) z2[i] = make_ulonglong2(0ull, 0ull);
        } else {
            for (int i = tid; i < cnt; i += TL_CTA_THREADS) src[i] = 0ull;
            for (int i = tid; i < ntgt; i += TL_CTA_THREADS) oth[i] = 0ull;
        }
        if (tid == 0) a.ticket[cloud] = 0u, a.done[cloud] = 0u;
    }
    HP_TRACE(tr0 + 8);
    if (flag_s[1]) asm volatile("griddepcontrol.wait;" ::: "memory");  // the tail grid does not complete before the ring grid
}

// ---- host side ---------------------------------------------------------------------------------------
// rounds per CTA: 1 unless the grid is many waves deep (then the row staging is amortised over more columns)
static int rf_rounds_per_cta(int b, int n, int m, int rows_per_cta) {
    const long long rowchunks = (n + rows_per_cta - 1) / rows_per_cta, rounds = (m + RF_COLS - 1) / RF_COLS;
    const long long want = (long long)sm_count() * 64 * (1024 / rows_per_cta);
    int R = 1;
    while (R < 8 && (long long)b * rowchunks * ((rounds + 2 * R - 1) / (2 * R)) >= want) R *= 2;
    return R;
}

// The ring kernel and the tail kernel share SMs while the tail runs under the ring kernel's last wave.  An SM cannot change
// its L1 / shared-memory split while CTAs are resident, so both kernels ask for the SAME (maximum shared memory) carve-out;
// otherwise a tail CTA can only start on an SM that has drained completely (measured: tails started 15 us late).
struct CarveCache {
    bool done[64] = {false};
};
template <typename K>
static cudaError_t ensure_max_carveout(K kern, CarveCache &cache) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && cache.done[dev]) return cudaSuccess;
#ifdef HP_BENCH_BUILD
    {
        const char *env = getenv("HP_NO_CARVEOUT");  // A/B: leave the driver's default L1 / shared-memory split
        if (env && atoi(env)) return cudaSuccess;
    }
#endif
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess && dev >= 0 && dev < 64) cache.done[dev] = true;
    return e;
}

template <int WARPS, int MINB, int DBG = 0>
static int ring_launch(const RingNNArgs &a, long long grid, cudaStream_t stream) {
    static CarveCache carve;
    HP_CUDA(ensure_max_carveout(nn_ring_kernel<WARPS, MINB, DBG>, carve));
    nn_ring_kernel<WARPS, MINB, DBG><<<(unsigned)grid, WARPS * 32, 0, stream>>>(a);
    return HP_OK;
}

// Per-cloud tickets (tails of early clouds run under the ring kernel's last wave, at the price of a fence at the end of every
// ring CTA) or whole-grid wait (no fence; every tail starts when the ring grid has completed).  The bench library can switch
// with HP_TAIL_TICKETS=0|1; the product build uses the constant.
#ifndef HP_TAIL_TICKETS_DEFAULT
#define HP_TAIL_TICKETS_DEFAULT 1
#endif
static int tail_tickets() {
#ifdef HP_BENCH_BUILD
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("HP_TAIL_TICKETS");
        v = e ? (atoi(e) != 0) : HP_TAIL_TICKETS_DEFAULT;
    }
    return v;
#else
    return HP_TAIL_TICKETS_DEFAULT;
#endif
}

// warps (x 256 rows) per ring CTA: the product build's constant; the bench library can override it with HP_RING_WARPS=1|2|4
static int ring_warps() {
#ifdef HP_BENCH_BUILD
    static int w = -1;
    if (w < 0) {
        const char *e = getenv("HP_RING_WARPS");
        w = e ? atoi(e) : RF_DEFAULT_WARPS;
        if (w != 1 && w != 2 && w != 4) w = RF_DEFAULT_WARPS;
    }
    return w;
#else
    return RF_DEFAULT_WARPS;
#endif
}

struct RFLayout {
    int chunks1, chunks2;
    size_t off_counters, off_ticket, off_done, off_losspart, off_rowkey, off_colkey, total;
};
static RFLayout rf_layout(int b, int n, int m) {
    RFLayout L;
    L.chunks1 = (n + LS_CHUNK - 1) / LS_CHUNK, L.chunks2 = (m + LS_CHUNK - 1) / LS_CHUNK;
    L.off_counters = 0;
    L.off_ticket = 16;
    L.off_done = L.off_ticket + sizeof(unsigned int) * (size_t)b;
    L.off_losspart = (L.off_done + sizeof(unsigned int) * (size_t)b + 15) & ~(size_t)15;
    L.off_rowkey = (L.off_losspart + (size_t)b * (L.chunks1 + L.chunks2) * sizeof(float) + 15) & ~(size_t)15;
    L.off_colkey = L.off_rowkey + (size_t)b * n * sizeof(u64);
    L.total = L.off_colkey + (size_t)b * m * sizeof(u64);
    return L;
}

size_t nn_ring_workspace_bytes(int b, int n, int m) {
    if (b <= 0 || n <= 0 || m <= 0) return 16;
    return rf_layout(b, n, m).total;
}

// Inputs must be finite with |coordinate| < 1e15 (padding points sit at +-1e18).  The workspace must be all zero on
// entry (every byte: the key arrays move with the shape) and is all zero again when the kernels of a call have run.
static int ceil_log2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}
// the in-kernel inverse needs both sort buffers in shared memory and the composite in 31 bits
bool nn_ring_inverse_supported(int n, int m) {
    if (n <= 0 || m <= 0) return false;
    const int big = n > m ? n : m;
    return big <= 32768;  // composite (target << shift | source) in 30 bits, sort buffer <= 128 KB
}

bool nn_ring_step_supported(int n, int m) { return n > 0 && m > 0 && rf_tail_smem_bytes(n, m) != 0; }

enum RingMode { RING_UNPACK = 0, RING_STEP = 1, RING_ONLY = 2 };

// The unpack kernel follows the ring kernel programmatically dependent (its CTAs become resident while the ring grid drains and
// start from griddepcontrol.wait the moment it completes).
static cudaError_t launch_unpack(void (*kern)(const RingNNArgs), unsigned grid, size_t smem, cudaStream_t stream, const RingNNArgs &a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(RF_MERGE_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs, cfg.numAttrs = 1;
#ifdef HP_BENCH_BUILD
    {
        const char *e = getenv("HP_NO_PDL");
        if (e && atoi(e)) cfg.numAttrs = 0;
    }
#endif
    return cudaLaunchKernelEx(&cfg, kern, a);
}

static int nn_ring_launch_impl(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1, int *idx1, float *dist2,
                               int *idx2, float *loss, int *inv1, int *inv2, const float *step_g, float *step_grad1,
                               float *step_grad2, void *workspace, cudaStream_t stream, RingMode mode);

int nn_ring_forward_launch(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1, int *idx1, float *dist2,
                           int *idx2, float *loss, int *inv1, int *inv2, void *workspace, cudaStream_t stream) {
    return nn_ring_launch_impl(b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, loss, inv1, inv2, nullptr, nullptr, nullptr, workspace,
                               stream, RING_UNPACK);
}

#ifdef HP_BENCH_BUILD
int nn_ring_set_trace(void *dev_ptr) {
    g_trace_host = reinterpret_cast<unsigned long long *>(dev_ptr);
    return HP_OK;
}
// bench library only (roofline of the dominant kernel): the ring kernel alone; leaves keys and tickets in the workspace
int nn_ring_only_launch(int b, int n, const float *xyz1, int m, const float *xyz2, void *workspace, cudaStream_t stream) {
    return nn_ring_launch_impl(b, n, xyz1, m, xyz2, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                               nullptr, workspace, stream, RING_ONLY);
}
#endif

// forward + backward of the fused loss in two kernels (ring + sectioned tail); requires nn_ring_step_supported(n, m)
int nn_ring_step_launch(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_loss, float *dist1, int *idx1,
                        float *dist2, int *idx2, float *loss, float *grad1, float *grad2, void *workspace, cudaStream_t stream) {
    return nn_ring_launch_impl(b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, loss, nullptr, nullptr, grad_loss, grad1, grad2,
                               workspace, stream, RING_STEP);
}

static int nn_ring_launch_impl(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1, int *idx1, float *dist2,
                               int *idx2, float *loss, int *inv1, int *inv2, const float *step_g, float *step_grad1,
                               float *step_grad2, void *workspace, cudaStream_t stream, RingMode mode) {
    RingNNArgs a = {};
    const RFLayout L = rf_layout(b, n, m);
    unsigned char *ws = reinterpret_cast<unsigned char *>(workspace);
    a.set1 = xyz1, a.set2 = xyz2, a.b = b, a.n = n, a.m = m;
    const int warps = ring_warps(), rows_per_cta = warps * RF_WROWS;
    a.tickets = (mode == RING_STEP && tail_tickets()) ? 1 : 0;
#ifdef HP_BENCH_BUILD
    a.trace = g_trace_host;
    {
        const char *env = getenv("HP_RING_ONLY_TICKETS");  // A/B: what the ticket arrival costs the ring kernel alone
        if (mode == RING_ONLY && env && atoi(env)) a.tickets = 1;
    }
#endif
    // whole-grid mode: every tail CTA should be resident when the ring grid completes -> four per SM: sources through L1
    a.stage_sources = a.tickets ? 1 : 0;
    a.R = rf_rounds_per_cta(b, n, m, rows_per_cta);
    a.rowchunks = (n + rows_per_cta - 1) / rows_per_cta;
    a.colchunks = ((m + RF_COLS - 1) / RF_COLS + a.R - 1) / a.R;
    a.counters = reinterpret_cast<unsigned int *>(ws + L.off_counters);
    a.ticket = reinterpret_cast<unsigned int *>(ws + L.off_ticket);
    a.done = reinterpret_cast<unsigned int *>(ws + L.off_done);
    a.losspart = reinterpret_cast<float *>(ws + L.off_losspart);
    a.rowkey = reinterpret_cast<u64 *>(ws + L.off_rowkey);
    a.colkey = reinterpret_cast<u64 *>(ws + L.off_colkey);
    a.dist1 = dist1, a.dist2 = dist2, a.idx1 = idx1, a.idx2 = idx2, a.loss = loss;
    const long long grid = (long long)b * a.rowchunks * a.colchunks;
    const long long ugrid = (long long)2 * b;
    HP_REQUIRE(grid <= 0x7fffffffLL && ugrid <= 0x7fffffffLL, "nn ring forward: grid too large (%lld CTAs)", grid);
    int rc_launch = HP_OK;
#ifdef HP_BENCH_BUILD
    static int variant = -1;
    if (variant < 0) {
        const char *e = getenv("HP_RING_VARIANT");
        variant = e ? atoi(e) : 0;
    }
    if (variant == 1) rc_launch = ring_launch<4, 5>(a, grid, stream);
    else if (variant == 2) rc_launch = ring_launch<4, 6>(a, grid, stream);
    else if (variant == 3) rc_launch = ring_launch<4, 3>(a, grid, stream);
    else if (variant == 10) rc_launch = ring_launch<4, 4, 1>(a, grid, stream);
    else if (variant == 11) rc_launch = ring_launch<4, 4, 2>(a, grid, stream);
    else if (variant == 12) rc_launch = ring_launch<4, 4, 3>(a, grid, stream);
    else if (variant == 20 && mode == RING_ONLY) rc_launch = ring_launch<4, 4, 4>(a, grid, stream);  // no tickets: never with a tail
    else
#endif
    if (warps == 1) rc_launch = ring_launch<1, 16>(a, grid, stream);
    else if (warps == 2) rc_launch = ring_launch<2, 8>(a, grid, stream);
    else rc_launch = ring_launch<4, 4>(a, grid, stream);
    if (rc_launch != HP_OK) return rc_launch;
    HP_LAUNCH_CHECK("nn_ring_kernel");
    if (mode == RING_ONLY) return HP_OK;  // measurement helper: keys and tickets stay in the workspace
    if (mode == RING_STEP) {
        const size_t smem = rf_tail_smem_bytes(n, m, a.stage_sources != 0);
        HP_REQUIRE(smem != 0 && loss != nullptr, "nn ring step: clouds too large for the fused tail (n=%d m=%d) or no loss output", n, m);
        RingTailArgs f;
        f.g = step_g, f.grad1 = step_grad1, f.grad2 = step_grad2;
        f.sec1 = (m + TL_SEC - 1) / TL_SEC, f.sec2 = (n + TL_SEC - 1) / TL_SEC;
        f.chunks1 = L.chunks1, f.chunks2 = L.chunks2;
        f.expected = a.tickets ? (unsigned)(a.rowchunks * a.colchunks) : 0u;
        const long long tgrid = (long long)b * (f.sec1 + f.sec2);
        HP_REQUIRE(tgrid <= 0x7fffffffLL, "nn ring step: grid too large (%lld CTAs)", tgrid);
        const bool small = (n > m ? n : m) <= 8 * TL_THREADS;
        static SmemAttrCache attr8, attr16;
        static CarveCache carve8, carve16;
        if (small) {
            HP_CUDA(ensure_dynamic_smem(nn_ring_tail_kernel<8>, smem, attr8));
            HP_CUDA(ensure_max_carveout(nn_ring_tail_kernel<8>, carve8));
        } else {
            HP_CUDA(ensure_dynamic_smem(nn_ring_tail_kernel<16>, smem, attr16));
            HP_CUDA(ensure_max_carveout(nn_ring_tail_kernel<16>, carve16));
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)tgrid), cfg.blockDim = dim3(TL_CTA_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
        cudaLaunchAttribute attrs[1];
        attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attrs, cfg.numAttrs = 1;
#ifdef HP_BENCH_BUILD
        {
            const char *e = getenv("HP_NO_PDL");
            if (e && atoi(e)) cfg.numAttrs = 0;
        }
#endif
        if (small) HP_CUDA(cudaLaunchKernelEx(&cfg, nn_ring_tail_kernel<8>, a, f));
        else HP_CUDA(cudaLaunchKernelEx(&cfg, nn_ring_tail_kernel<16>, a, f));
    } else if (inv1 != nullptr && inv2 != nullptr) {
        HP_REQUIRE(nn_ring_inverse_supported(n, m), "nn ring forward: clouds too large for the in-kernel inverse (n=%d m=%d)", n, m);
        a.inv1 = inv1, a.inv2 = inv2;
        const int big = n > m ? n : m;
        a.sort_n = 1 << ceil_log2(big);
        a.sort_shift = ceil_log2(big);
        // counting sort (keys | count | cursor | pbuf = 4*big ints) with radix fallback + its 32 x 64 rank table (<= 8192 points);
        // larger clouds: one bitonic buffer of sort_n uints
        a.inv_fast = (big <= 8192) ? 1 : 0;
        const size_t smem = a.inv_fast ? ((size_t)4 * big + RX_WARPS * RX_BINS + RX_BINS) * sizeof(unsigned int)
                                       : (size_t)a.sort_n * sizeof(unsigned int);
        static SmemAttrCache attr;
        if (smem > 40 * 1024) HP_CUDA(ensure_dynamic_smem(nn_ring_unpack_kernel<true>, smem, attr));
        HP_CUDA(launch_unpack(nn_ring_unpack_kernel<true>, (unsigned)ugrid, smem, stream, a));
    } else {
        HP_CUDA(launch_unpack(nn_ring_unpack_kernel<false>, (unsigned)ugrid, 0, stream, a));
    }
    HP_LAUNCH_CHECK("nn_ring_unpack_kernel / nn_ring_tail_kernel");
    return HP_OK;
}


// ---- backward as a pure gather over the inverse maps produced by the forward --------------------------------
//   grad_a[i] = 2 g [ (a_i - b_idx1[i]) + sum_{k in bucket2(i), ascending} (a_i - b_k) ]      (nndistance.cu:143-151)
//   grad_b[k] = 2 g [ (b_k - a_idx2[k]) + sum_{i in bucket1(k), ascending} (b_k - a_i) ]
// One thread per point, all clouds and both sides in one launch; same summation order as nn_grad_kernel.
struct RingGradArgs {
    const float *set1, *set2;
    const int *idx1, *idx2, *inv1, *inv2;
    const float *g;          // one device scalar (the loss gradient), or nullptr when per-point gradients are given
    const float *gd1, *gd2;  // per-point upstream gradients [b,n], [b,m] (nn_distance), or nullptr
    float *grad1, *grad2;
    int b, n, m;
};

// Buckets above GATHER_COOP entries (skewed assignments: early in training most ground-truth points map to a few
// reconstructed points) are summed by the whole warp: lane l takes entries l, l+32, ... in ascending order and the 32
// partial sums are folded by a fixed shuffle tree -- still deterministic, and no thread walks thousands of entries alone.
constexpr int GATHER_COOP = 32;

__global__ void __launch_bounds__(256) nn_grad_gather_kernel(const RingGradArgs a) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total1 = (long long)a.b * a.n, total = total1 + (long long)a.b * a.m;
    const int lane = threadIdx.x & 31;
    const bool valid = t < total;
    const bool side2 = valid && t >= total1;
    const long long u = side2 ? t - total1 : (valid ? t : 0);
    const int np = side2 ? a.m : a.n, no = side2 ? a.n : a.m;
    const int cloud = (int)(u / np), i = (int)(u - (long long)cloud * np);
    const float *__restrict__ P = (side2 ? a.set2 : a.set1) + (size_t)cloud * np * 3;
    const float *__restrict__ O = (side2 ? a.set1 : a.set2) + (size_t)cloud * no * 3;
    const int *__restrict__ idx_own = (side2 ? a.idx2 : a.idx1) + (size_t)cloud * np;
    // buckets of the OTHER side's index map: inv of the other direction = [perm[no] | begin[np] | end[np]]
    const int *__restrict__ perm = (side2 ? a.inv1 : a.inv2) + (size_t)cloud * (no + 2 * np);
    const float *__restrict__ g_own = a.g ? nullptr : (side2 ? a.gd2 : a.gd1) + (size_t)cloud * np;
    const float *__restrict__ g_oth = a.g ? nullptr : (side2 ? a.gd1 : a.gd2) + (size_t)cloud * no;
    const float gs = a.g ? __ldg(a.g) * 2.f : 0.f;
    int pb = 0, pe = 0;
    float px = 0.f, py = 0.f, pz = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
    if (valid) {
        pb = __ldg(perm + no + i), pe = __ldg(perm + no + np + i);
        const float g2 = a.g ? gs : __ldg(g_own + i) * 2.f;
        px = __ldg(P + (size_t)i * 3 + 0), py = __ldg(P + (size_t)i * 3 + 1), pz = __ldg(P + (size_t)i * 3 + 2);
        const int j = min(max(__ldg(idx_own + i), 0), no - 1);
        ax = g2 * (px - __ldg(O + (size_t)j * 3 + 0));
        ay = g2 * (py - __ldg(O + (size_t)j * 3 + 1));
        az = g2 * (pz - __ldg(O + (size_t)j * 3 + 2));
    }
    const bool big = valid && (pe - pb) > GATHER_COOP;
    if (valid && !big) {
        for (int p = pb; p < pe; ++p) {  // ascending source index: fixed summation order
            const int k = __ldg(perm + p);
            const float gk = a.g ? gs : __ldg(g_oth + k) * 2.f;
            ax += -(gk * (__ldg(O + (size_t)k * 3 + 0) - px));
            ay += -(gk * (__ldg(O + (size_t)k * 3 + 1) - py));
            az += -(gk * (__ldg(O + (size_t)k * 3 + 2) - pz));
        }
    }
    // warp-cooperative pass over the big buckets of this warp's 32 points (rare; uniform loop over the ballot)
    unsigned todo = __ballot_sync(0xffffffffu, big);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int b0 = __shfl_sync(0xffffffffu, pb, src), b1 = __shfl_sync(0xffffffffu, pe, src);
        const float qx = __shfl_sync(0xffffffffu, px, src), qy = __shfl_sync(0xffffffffu, py, src), qz = __shfl_sync(0xffffffffu, pz, src);
        // all lanes of a warp may straddle a cloud / side boundary: take the owner's arrays
        const unsigned long long permv = __shfl_sync(0xffffffffu, (unsigned long long)perm, src);
        const unsigned long long Ov = __shfl_sync(0xffffffffu, (unsigned long long)O, src);
        const unsigned long long gv = __shfl_sync(0xffffffffu, (unsigned long long)g_oth, src);
        const int *__restrict__ perm_s = reinterpret_cast<const int *>(permv);
        const float *__restrict__ O_s = reinterpret_cast<const float *>(Ov);
        const float *__restrict__ g_s = reinterpret_cast<const float *>(gv);
        float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll 4
        for (int p = b0 + lane; p < b1; p += 32) {  // unrolled: four dependent (perm -> point) load chains in flight
            const int k = __ldg(perm_s + p);
            const float gk = a.g ? gs : __ldg(g_s + k) * 2.f;
            sx += -(gk * (__ldg(O_s + (size_t)k * 3 + 0) - qx));
            sy += -(gk * (__ldg(O_s + (size_t)k * 3 + 1) - qy));
            sz += -(gk * (__ldg(O_s + (size_t)k * 3 + 2) - qz));
        }
        sx = warp_sum(sx), sy = warp_sum(sy), sz = warp_sum(sz);
        if (lane == src) ax += sx, ay += sy, az += sz;
    }
    if (valid) {
        float *G = (side2 ? a.grad2 : a.grad1) + ((size_t)cloud * np + i) * 3;
        G[0] = ax, G[1] = ay, G[2] = az;
    }
}

int nn_ring_backward_launch(int b, int n, const float *xyz1, int m, const float *xyz2, const int *idx1, const int *idx2,
                            const int *inv1, const int *inv2, const float *grad_loss, const float *grad_dist1,
                            const float *grad_dist2, float *grad1, float *grad2, cudaStream_t stream) {
    RingGradArgs a;
    a.set1 = xyz1, a.set2 = xyz2, a.idx1 = idx1, a.idx2 = idx2, a.inv1 = inv1, a.inv2 = inv2, a.g = grad_loss;
    a.gd1 = grad_dist1, a.gd2 = grad_dist2;
    a.grad1 = grad1, a.grad2 = grad2, a.b = b, a.n = n, a.m = m;
    const long long total = (long long)b * ((long long)n + m);
    const long long grid = (total + 255) / 256;
    HP_REQUIRE(grid <= 0x7fffffffLL, "nn ring backward: grid too large");
    nn_grad_gather_kernel<<<(unsigned)grid, 256, 0, stream>>>(a);
    HP_LAUNCH_CHECK("nn_grad_gather_kernel");
    return HP_OK;
}

}  // namespace hp
