// pairwise_cd.cu -- all-pairs Chamfer cloud-distance matrix for the set-vs-set metrics.
//
// Replaces the CD half of _pairwise_EMD_CD_ / dist_chamfer (utils/metrics.py:78-83,121-158): for every
// cloud pair (r, s)   cd[r,s] = mean_i min_j |a_i - b_j|^2 + mean_j min_i |a_i - b_j|^2.
// The reference expands one cloud to the chunk's batch, materialises P[B,N,M] with three bmm's and
// takes two min reductions.  Here one CTA owns one cloud pair and evaluates every UNORDERED point pair
// exactly once, feeding both minima, entirely in registers:
//
//   "warp ring": lane L keeps RQ=8 points of cloud r (rows) and RC=4 points of cloud s (columns) in
//   registers.  32 times per round: evaluate the RQ x RC distances (FADD2/FMUL2/FFMA2, the library's
//   one exact association), fold them into the lane's 8 running row minima and into the 4 running
//   column minima that TRAVEL WITH the column points, then rotate the column points and their minima to
//   the next lane with SHFL.  After 32 rotations every column point has met all 256 row points of the
//   warp and is back home.  No shared-memory traffic in the hot loop: the kernel is bound by FP32
//   issue (6 FP32 ops per unordered pair + ~1.5 min/shuffle ops), not by LDS bandwidth.
//
// Column minima of the CTA's warps are merged with integer atomicMin on the float bit patterns
// (d >= 0, so uint order == float order; integer atomics are order-independent => deterministic).
// Cloud s is staged in shared memory once per CTA with a 1-D bulk-TMA copy.
#include "common.cuh"

namespace hp {

constexpr int RING_RQ = 8;                  // row points per lane
constexpr int RING_RC = 4;                  // column points per lane
constexpr int RING_ROWS = 32 * RING_RQ;     // 256 rows per warp block
constexpr int RING_COLS = 32 * RING_RC;     // 128 columns per round
constexpr int PCD_WARPS = 8;
constexpr int PCD_CHUNK = 2048;             // column points staged per shared-memory chunk (24 KB)
constexpr float RING_PAD_ROW = 1.0e18f;     // padding points: farther from everything than any real pair,
constexpr float RING_PAD_COL = -1.0e18f;    // yet finite (no inf-inf NaNs)

// Column points live as packed fp32x2 PAIRS (columns 0|1 and 2|3) so that the loop-carried values are
// aligned 64-bit registers: the SHFLs rotate them in place and the FADD2/FFMA2 consume them without a
// single register move.
struct RingCols {
    f32x2 x01, x23, y01, y23, z01, z23;
    float mn[RING_RC];
};

__device__ __forceinline__ f32x2 shfl2(f32x2 v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// One round: 32 rotations of the column registers through the warp.
__device__ __forceinline__ void ring_round(const float (&qx)[RING_RQ], const float (&qy)[RING_RQ],
                                           const float (&qz)[RING_RQ], float (&rowbest)[RING_RQ], RingCols &c,
                                           const int lane) {
    const int src = (lane + 1) & 31;
#pragma unroll 1
    for (int rot = 0; rot < 32; ++rot) {
#pragma unroll
        for (int j = 0; j < RING_RQ; j += 2) {
            float d[2][4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const f32x2 px = pack2(qx[j + u], qx[j + u]), py = pack2(qy[j + u], qy[j + u]), pz = pack2(qz[j + u], qz[j + u]);
                unpack2(sqdist_exact2(px, py, pz, c.x01, c.y01, c.z01), d[u][0], d[u][1]);
                unpack2(sqdist_exact2(px, py, pz, c.x23, c.y23, c.z23), d[u][2], d[u][3]);
                rowbest[j + u] = min3(rowbest[j + u], d[u][0], d[u][1]);
                rowbest[j + u] = min3(rowbest[j + u], d[u][2], d[u][3]);
            }
#pragma unroll
            for (int i = 0; i < RING_RC; ++i) c.mn[i] = min3(c.mn[i], d[0][i], d[1][i]);
        }
        c.x01 = shfl2(c.x01, src), c.x23 = shfl2(c.x23, src);
        c.y01 = shfl2(c.y01, src), c.y23 = shfl2(c.y23, src);
        c.z01 = shfl2(c.z01, src), c.z23 = shfl2(c.z23, src);
#pragma unroll
        for (int i = 0; i < RING_RC; ++i) c.mn[i] = __shfl_sync(0xffffffffu, c.mn[i], src);
    }
}

__global__ void __launch_bounds__(PCD_WARPS * 32, 2)
    pairwise_cd_kernel(int nb, int n, int m, const float *__restrict__ first, const float *__restrict__ second,
                       int row_begin, const int *__restrict__ pair_r, const int *__restrict__ pair_s, float *__restrict__ cd) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *stage = reinterpret_cast<float *>(smem_raw);                        // [PCD_CHUNK*3] cloud-s chunk, AoS
    unsigned int *colmin = reinterpret_cast<unsigned int *>(stage + PCD_CHUNK * 3);  // [m] float bits
    __shared__ __align__(8) uint64_t mbar;
    __shared__ float part[2][PCD_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // cloud pair of this CTA: entry blockIdx.x of an explicit pair list, or (row, column) of a row block of the full matrix
    const int r = pair_r ? __ldg(pair_r + blockIdx.x) : row_begin + (int)(blockIdx.x / nb);
    const int s = pair_s ? __ldg(pair_s + blockIdx.x) : (int)(blockIdx.x % nb);
    const float *__restrict__ A = first + (size_t)r * n * 3;
    const float *__restrict__ Bp = second + (size_t)s * m * 3;

    if (tid == 0) mbar_init(&mbar, 1);
    for (int i = tid; i < m; i += PCD_WARPS * 32) colmin[i] = 0x7f800000u;
    __syncthreads();

    float rowsum = 0.f;
    uint32_t phase = 0;
    for (int qs = 0; qs < n; qs += PCD_WARPS * RING_ROWS) {
        const int q0 = qs + warp * RING_ROWS + lane * RING_RQ;  // this lane's 8 consecutive rows
        const bool warp_active = qs + warp * RING_ROWS < n;     // warp-uniform
        float qx[RING_RQ], qy[RING_RQ], qz[RING_RQ], rowbest[RING_RQ];
#pragma unroll
        for (int j = 0; j < RING_RQ; ++j) {
            const int q = q0 + j;
            qx[j] = qy[j] = qz[j] = RING_PAD_ROW;
            if (q < n) qx[j] = __ldg(A + (size_t)q * 3 + 0), qy[j] = __ldg(A + (size_t)q * 3 + 1), qz[j] = __ldg(A + (size_t)q * 3 + 2);
            rowbest[j] = __int_as_float(0x7f800000);
        }
        for (int ch = 0; ch < m; ch += PCD_CHUNK) {
            const int cnt = min(PCD_CHUNK, m - ch);
            const float *src = Bp + (size_t)ch * 3;
            __syncthreads();  // previous chunk fully consumed
            const bool tma_ok = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
            const uint32_t bulk_bytes = tma_ok ? ((uint32_t)(cnt * 12) & ~15u) : 0u;
            if (bulk_bytes && tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(&mbar, bulk_bytes);
                bulk_g2s(stage, src, bulk_bytes, &mbar);
            }
            for (int i = (int)(bulk_bytes / 4) + tid; i < cnt * 3; i += PCD_WARPS * 32) stage[i] = __ldg(src + i);
            if (bulk_bytes) {
                mbar_wait(&mbar, phase);
                phase ^= 1;
            }
            __syncthreads();
            if (warp_active) {
                for (int g = 0; g < cnt; g += RING_COLS) {
                    RingCols c;
                    const int c0 = g + lane * RING_RC;  // this lane's 4 consecutive columns within the chunk
                    if (c0 + RING_RC <= cnt) {
                        const float4 v0 = *reinterpret_cast<const float4 *>(stage + (size_t)c0 * 3);
                        const float4 v1 = *reinterpret_cast<const float4 *>(stage + (size_t)c0 * 3 + 4);
                        const float4 v2 = *reinterpret_cast<const float4 *>(stage + (size_t)c0 * 3 + 8);
                        c.x01 = pack2(v0.x, v0.w), c.y01 = pack2(v0.y, v1.x), c.z01 = pack2(v0.z, v1.y);
                        c.x23 = pack2(v1.z, v2.y), c.y23 = pack2(v1.w, v2.z), c.z23 = pack2(v2.x, v2.w);
                    } else {
                        float tx[RING_RC], ty[RING_RC], tz[RING_RC];
#pragma unroll
                        for (int i = 0; i < RING_RC; ++i) {
                            tx[i] = ty[i] = tz[i] = RING_PAD_COL;
                            if (c0 + i < cnt) tx[i] = stage[(c0 + i) * 3 + 0], ty[i] = stage[(c0 + i) * 3 + 1], tz[i] = stage[(c0 + i) * 3 + 2];
                        }
                        c.x01 = pack2(tx[0], tx[1]), c.y01 = pack2(ty[0], ty[1]), c.z01 = pack2(tz[0], tz[1]);
                        c.x23 = pack2(tx[2], tx[3]), c.y23 = pack2(ty[2], ty[3]), c.z23 = pack2(tz[2], tz[3]);
                    }
#pragma unroll
                    for (int i = 0; i < RING_RC; ++i) c.mn[i] = __int_as_float(0x7f800000);
                    ring_round(qx, qy, qz, rowbest, c, lane);
#pragma unroll
                    for (int i = 0; i < RING_RC; ++i)
                        if (c0 + i < cnt) atomicMin(&colmin[ch + c0 + i], __float_as_uint(c.mn[i]));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < RING_RQ; ++j)
            if (q0 + j < n) rowsum += rowbest[j];
    }
    __syncthreads();
    float colsum = 0.f;
    for (int i = tid; i < m; i += PCD_WARPS * 32) colsum += __uint_as_float(colmin[i]);
    rowsum = warp_sum(rowsum);
    colsum = warp_sum(colsum);
    if (lane == 0) part[0][warp] = rowsum, part[1][warp] = colsum;
    __syncthreads();
    if (tid == 0) {
        float rs = 0.f, cs = 0.f;
#pragma unroll
        for (int w = 0; w < PCD_WARPS; ++w) rs += part[0][w], cs += part[1][w];
        cd[blockIdx.x] = rs / (float)n + cs / (float)m;  // dl.mean(1) + dr.mean(1), utils/metrics.py:145
    }
}

}  // namespace hp

using namespace hp;

// one cache for both entry points: the attribute belongs to the kernel, not to the caller
static SmemAttrCache g_pcd_smem_attr;

extern "C" int hp_pairwise_cd(int na, int nb, int n, int m, const float *first, const float *second, int row_begin,
                              int row_end, float *cd, void *stream) {
    HP_REQUIRE(na >= 0 && nb >= 0 && n >= 0 && m >= 0, "hp_pairwise_cd: negative size");
    HP_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= na, "hp_pairwise_cd: bad row range [%d,%d) of %d", row_begin,
               row_end, na);
    const long long pairs = (long long)(row_end - row_begin) * nb;
    if (pairs == 0) return HP_OK;
    HP_REQUIRE(n > 0 && m > 0, "hp_pairwise_cd: empty clouds (n=%d m=%d)", n, m);
    HP_REQUIRE(first && second && cd, "hp_pairwise_cd: null pointer");
    HP_REQUIRE(pairs <= 0x7fffffffLL, "hp_pairwise_cd: %lld cloud pairs in one call; split the row range", pairs);
    const size_t smem = (size_t)PCD_CHUNK * 3 * sizeof(float) + (size_t)m * sizeof(unsigned int);
    if (smem > 200 * 1024) {
        set_error("hp_pairwise_cd: m=%d column points need %zu bytes of shared memory (limit 200 KB)", m, smem);
        return HP_ERR_UNSUPPORTED;
    }
    HP_CUDA(ensure_dynamic_smem(pairwise_cd_kernel, smem, g_pcd_smem_attr));
    pairwise_cd_kernel<<<(unsigned)pairs, PCD_WARPS * 32, smem, (cudaStream_t)stream>>>(nb, n, m, first, second, row_begin, nullptr,
                                                                                       nullptr, cd);
    HP_LAUNCH_CHECK("pairwise_cd_kernel");
    return HP_OK;
}

extern "C" int hp_pairwise_cd_pairs(long long npairs, int n, int m, const float *first, const float *second, const int *pair_r,
                                    const int *pair_s, float *cd, void *stream) {
    HP_REQUIRE(npairs >= 0 && n >= 0 && m >= 0, "hp_pairwise_cd_pairs: negative size");
    if (npairs == 0) return HP_OK;
    HP_REQUIRE(n > 0 && m > 0, "hp_pairwise_cd_pairs: empty clouds (n=%d m=%d)", n, m);
    HP_REQUIRE(first && second && pair_r && pair_s && cd, "hp_pairwise_cd_pairs: null pointer");
    HP_REQUIRE(npairs <= 0x7fffffffLL, "hp_pairwise_cd_pairs: %lld cloud pairs in one call; split the list", npairs);
    const size_t smem = (size_t)PCD_CHUNK * 3 * sizeof(float) + (size_t)m * sizeof(unsigned int);
    if (smem > 200 * 1024) {
        set_error("hp_pairwise_cd_pairs: m=%d column points need %zu bytes of shared memory (limit 200 KB)", m, smem);
        return HP_ERR_UNSUPPORTED;
    }
    HP_CUDA(ensure_dynamic_smem(pairwise_cd_kernel, smem, g_pcd_smem_attr));
    pairwise_cd_kernel<<<(unsigned)npairs, PCD_WARPS * 32, smem, (cudaStream_t)stream>>>(1, n, m, first, second, 0, pair_r, pair_s, cd);
    HP_LAUNCH_CHECK("pairwise_cd_kernel (pair list)");
    return HP_OK;
}
