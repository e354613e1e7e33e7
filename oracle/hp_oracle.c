/*
 * oracle/hp_oracle.c -- TEST INFRASTRUCTURE ONLY. Never linked, imported or called by the
 * product path (3d-point-clouds-autocomplete_b200/); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * A plain-C, single-source CPU restatement of the arithmetic of the reference's CUDA
 * structural-loss backend (gmum/3d-point-clouds-autocomplete,
 * utils/pytorch_structural_losses/{nndistance.cu, approxmatch.cu}).  Every function cites
 * the reference lines it follows.  It restates WHAT each kernel computes (operation order,
 * fp32 association, tie rules, epsilons), written as ordinary loops over "virtual threads";
 * it is not a copy of the CUDA sources.
 *
 * Pinning: the reference ships no tests, golden vectors or fixtures for this path
 * (SURVEY.md section 4 / 8c).  This oracle is therefore pinned differentially:
 *   - on the GPU box against the UNMODIFIED reference extension built by
 *     oracle/build_ref.sh into oracle/_ref/ (tests/test_gpu_vs_reference_ext.py), and
 *   - against tests/golden/ fixtures produced by that extension on a B200
 *     (tests/golden/make_golden_gpu.py) and by the reference's pure-torch modules on CPU
 *     (tests/golden/make_golden_cpu.py).
 *
 * fp32 association (verified in the sm_100 SASS of the reference build, nvcc 12.9):
 *     d = fma(dz, dz, fma(dx, dx, dy * dy))          dx = cand.x - query.x, ...
 * Compile with -ffp-contract=off so the C compiler adds no contractions of its own.
 *
 * Differences that cannot be restated bit-exactly on a CPU (documented, tolerance 1e-5):
 *   - __expf  = ex2.approx(x * log2e_f32)   -> here exp2f(x * 1.4426950216f)
 *   - rsqrtf  = MUFU.RSQ (approx)           -> here 1.0f / sqrtf(x)
 *   - float atomicAdd order in NmDistanceGradKernel is unspecified -> here ascending j.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define HP_EXPORT __attribute__((visibility("default")))

static inline float sqdist3(float qx, float qy, float qz, float cx, float cy, float cz) {
    /* nndistance.cu:28-31 / approxmatch.cu:85 as compiled: fma(dz,dz,fma(dx,dx,dy*dy)) */
    float dx = cx - qx, dy = cy - qy, dz = cz - qz;
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

/* ---------------------------------------------------------------------------------------
 * One direction of the nearest-neighbour search.
 * Follows NmDistanceKernel, nndistance.cu:8-130:
 *   - candidates are visited in tiles of 512 (nndistance.cu:9,12);
 *   - inside a tile the first candidate always wins, later ones only on strict '<'
 *     (nndistance.cu:32,42,52,62,112);
 *   - a later tile replaces the stored result only on strict '>' (nndistance.cu:122).
 *   => lowest candidate index among the exact minima.
 * ------------------------------------------------------------------------------------- */
static void nn_one_direction(int b, int n, const float *xyz, int m, const float *xyz2,
                             float *result, int *result_i) {
    const int tile = 512;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < b; ++i) {
        for (int k2 = 0; k2 < m; k2 += tile) {
            int end_k = (m < k2 + tile ? m : k2 + tile) - k2;
            const float *cand = xyz2 + ((size_t)i * m + k2) * 3;
            for (int j = 0; j < n; ++j) {
                const float *q = xyz + ((size_t)i * n + j) * 3;
                float best = 0.0f;
                int best_i = 0;
                for (int k = 0; k < end_k; ++k) {
                    float d = sqdist3(q[0], q[1], q[2], cand[k * 3 + 0], cand[k * 3 + 1], cand[k * 3 + 2]);
                    if (k == 0 || d < best) {
                        best = d;
                        best_i = k + k2;
                    }
                }
                size_t o = (size_t)i * n + j;
                if (k2 == 0 || result[o] > best) {
                    result[o] = best;
                    result_i[o] = best_i;
                }
            }
        }
    }
}

/* nndistance(), nndistance.cu:131-134: two launches, second with the roles swapped. */
HP_EXPORT void hp_oracle_nndistance(int b, int n, const float *xyz, int m, const float *xyz2,
                                    float *result, int *result_i, float *result2, int *result2_i) {
    nn_one_direction(b, n, xyz, m, xyz2, result, result_i);
    nn_one_direction(b, m, xyz2, n, xyz, result2, result2_i);
}

/* NmDistanceGradKernel, nndistance.cu:135-154 (one direction).  The reference accumulates
 * with float atomicAdd in an unspecified order; here ascending j. */
static void nn_grad_one_direction(int b, int n, const float *xyz1, int m, const float *xyz2,
                                  const float *grad_dist1, const int *idx1, float *grad_xyz1,
                                  float *grad_xyz2) {
    for (int i = 0; i < b; ++i) {
        for (int j = 0; j < n; ++j) {
            size_t o = (size_t)i * n + j;
            float x1 = xyz1[o * 3 + 0], y1 = xyz1[o * 3 + 1], z1 = xyz1[o * 3 + 2];
            int j2 = idx1[o];
            size_t o2 = (size_t)i * m + j2;
            float x2 = xyz2[o2 * 3 + 0], y2 = xyz2[o2 * 3 + 1], z2 = xyz2[o2 * 3 + 2];
            float g = grad_dist1[o] * 2;
            grad_xyz1[o * 3 + 0] += g * (x1 - x2);
            grad_xyz1[o * 3 + 1] += g * (y1 - y2);
            grad_xyz1[o * 3 + 2] += g * (z1 - z2);
            grad_xyz2[o2 * 3 + 0] += -(g * (x1 - x2));
            grad_xyz2[o2 * 3 + 1] += -(g * (y1 - y2));
            grad_xyz2[o2 * 3 + 2] += -(g * (z1 - z2));
        }
    }
}

/* nndistancegrad(), nndistance.cu:155-160: zero both gradients, then both directions. */
HP_EXPORT void hp_oracle_nndistancegrad(int b, int n, const float *xyz1, int m, const float *xyz2,
                                        const float *grad_dist1, const int *idx1,
                                        const float *grad_dist2, const int *idx2,
                                        float *grad_xyz1, float *grad_xyz2) {
    memset(grad_xyz1, 0, (size_t)b * n * 3 * sizeof(float));
    memset(grad_xyz2, 0, (size_t)b * m * 3 * sizeof(float));
    nn_grad_one_direction(b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_xyz1, grad_xyz2);
    nn_grad_one_direction(b, m, xyz2, n, xyz1, grad_dist2, idx2, grad_xyz2, grad_xyz1);
}

/* __expf(x) as the reference build evaluates it: ex2.approx(x * log2e) (approxmatch.cu:86,
 * 131,185; SASS: FMUL x,1.4426950216 ; MUFU.EX2 with the <-126 rescale).  exp2f stands in
 * for MUFU.EX2 (<= 2 ulp apart). */
static inline float fast_expf(float x) { return exp2f(x * 1.4426950216293334961f); }

/* ---------------------------------------------------------------------------------------
 * approxmatchkernel, approxmatch.cu:34-213.  One cloud pair at a time.
 *   match : [b][m][n]  (approxmatch.cu:186  match[i*n*m + l*n + k])
 *   temp  : [b][2(n+m)] = remainL[n] remainR[m] ratioL[n] ratioR[m]   (approxmatch.cu:35;
 *           the reference indexes temp by blockIdx.x, which equals the cloud index for
 *           b <= 32; this restatement gives each cloud its own slice).
 * Level schedule: j = 7 .. -1, level = -4^j (approxmatch.cu:55-59; the j == -2 branch is
 * dead).  multiL / multiR use integer division (approxmatch.cu:37-43).
 * Accumulation order inside every sum is ascending index, as each CUDA thread does it.
 * ------------------------------------------------------------------------------------- */
HP_EXPORT void hp_oracle_approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2,
                                     float *match, float *temp) {
    float multiL, multiR;
    if (n >= m) {
        multiL = 1;
        multiR = (float)(n / m);
    } else {
        multiL = (float)(m / n);
        multiR = 1;
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < b; ++i) {
        float *remainL = temp + (size_t)i * (n + m) * 2;
        float *remainR = remainL + n;
        float *ratioL = remainR + m;
        float *ratioR = ratioL + n;
        const float *p1 = xyz1 + (size_t)i * n * 3;
        const float *p2 = xyz2 + (size_t)i * m * 3;
        float *mt = match + (size_t)i * n * m;
        memset(mt, 0, (size_t)n * m * sizeof(float));
        for (int k = 0; k < n; ++k) remainL[k] = multiL;
        for (int l = 0; l < m; ++l) remainR[l] = multiR;
        for (int j = 7; j > -2; --j) {
            float level = -powf(4.0f, (float)j);
            /* pass 1, approxmatch.cu:60-93: ratioL[k] = remainL[k] / (1e-9 + sum_l e^{level d} remainR[l]) */
            for (int k = 0; k < n; ++k) {
                float suml = 1e-9f;
                for (int l = 0; l < m; ++l) {
                    float d = level * sqdist3(p1[k * 3], p1[k * 3 + 1], p1[k * 3 + 2],
                                              p2[l * 3], p2[l * 3 + 1], p2[l * 3 + 2]);
                    suml = fmaf(fast_expf(d), remainR[l], suml);
                }
                ratioL[k] = remainL[k] / suml;
            }
            /* pass 2, approxmatch.cu:109-142 */
            for (int l = 0; l < m; ++l) {
                float sumr = 0;
                for (int k = 0; k < n; ++k) {
                    float d = level * sqdist3(p2[l * 3], p2[l * 3 + 1], p2[l * 3 + 2],
                                              p1[k * 3], p1[k * 3 + 1], p1[k * 3 + 2]);
                    sumr = fmaf(fast_expf(d), ratioL[k], sumr);
                }
                sumr *= remainR[l];
                float consumption = fminf(remainR[l] / (sumr + 1e-9f), 1.0f);
                ratioR[l] = consumption * remainR[l];
                remainR[l] = fmaxf(0.0f, remainR[l] - sumr);
            }
            /* pass 3, approxmatch.cu:161-194 */
            for (int k = 0; k < n; ++k) {
                float suml = 0;
                float rl = ratioL[k];
                for (int l = 0; l < m; ++l) {
                    float d = level * sqdist3(p1[k * 3], p1[k * 3 + 1], p1[k * 3 + 2],
                                              p2[l * 3], p2[l * 3 + 1], p2[l * 3 + 2]);
                    /* as compiled (sm_100 SASS of the reference): t = rl*e, then BOTH updates are fma(t, ratioR, .) */
                    float t = rl * fast_expf(d);
                    mt[(size_t)l * n + k] = fmaf(t, ratioR[l], mt[(size_t)l * n + k]);
                    suml = fmaf(t, ratioR[l], suml);
                }
                remainL[k] = fmaxf(0.0f, remainL[k] - suml);
            }
        }
    }
}

/* matchcostkernel, approxmatch.cu:215-255: 512 virtual threads; thread t accumulates, for
 * each 256-candidate tile in order, its points j = t, t+512, ... ; then the pairwise tree
 * of approxmatch.cu:245-250. */
HP_EXPORT void hp_oracle_matchcost(int b, int n, int m, const float *xyz1, const float *xyz2,
                                   const float *match, float *out) {
    const int T = 512, Block = 256;
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < b; ++i) {
        float allsum[512];
        for (int t = 0; t < T; ++t) {
            float subsum = 0;
            for (int k0 = 0; k0 < m; k0 += Block) {
                int endk = m < k0 + Block ? m : k0 + Block;
                for (int j = t; j < n; j += T) {
                    const float *q = xyz1 + ((size_t)i * n + j) * 3;
                    for (int k = k0; k < endk; ++k) {
                        const float *c = xyz2 + ((size_t)i * m + k) * 3;
                        float d = sqrtf(sqdist3(q[0], q[1], q[2], c[0], c[1], c[2]));
                        subsum = fmaf(match[(size_t)i * n * m + (size_t)k * n + j], d, subsum);
                    }
                }
            }
            allsum[t] = subsum;
        }
        for (int j = 1; j < T; j <<= 1)
            for (int t = 0; t < T; ++t)
                if ((t & j) == 0 && t + j < T && (t & (j - 1)) == 0) allsum[t] += allsum[t + j];
        out[i] = allsum[0];
    }
}

/* matchcostgrad1kernel (approxmatch.cu:301-322) and matchcostgrad2kernel (:260-300). */
HP_EXPORT void hp_oracle_matchcostgrad(int b, int n, int m, const float *xyz1, const float *xyz2,
                                       const float *match, float *grad1, float *grad2) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < b; ++i) {
        for (int l = 0; l < n; ++l) {
            const float *q = xyz1 + ((size_t)i * n + l) * 3;
            float dx = 0, dy = 0, dz = 0;
            for (int k = 0; k < m; ++k) {
                const float *c = xyz2 + ((size_t)i * m + k) * 3;
                float ex = q[0] - c[0], ey = q[1] - c[1], ez = q[2] - c[2];
                float s = fmaf(ez, ez, fmaf(ex, ex, ey * ey));
                float d = match[(size_t)i * n * m + (size_t)k * n + l] * (1.0f / sqrtf(fmaxf(s, 1e-20f)));
                dx = fmaf(ex, d, dx);
                dy = fmaf(ey, d, dy);
                dz = fmaf(ez, d, dz);
            }
            grad1[((size_t)i * n + l) * 3 + 0] = dx;
            grad1[((size_t)i * n + l) * 3 + 1] = dy;
            grad1[((size_t)i * n + l) * 3 + 2] = dz;
        }
        const int T = 256;
        float sg[256 * 3];
        for (int k = 0; k < m; ++k) {
            const float *c = xyz2 + ((size_t)i * m + k) * 3;
            for (int t = 0; t < T; ++t) {
                float sx = 0, sy = 0, sz = 0;
                for (int j = t; j < n; j += T) {
                    const float *q = xyz1 + ((size_t)i * n + j) * 3;
                    float ex = c[0] - q[0], ey = c[1] - q[1], ez = c[2] - q[2];
                    float s = fmaf(ez, ez, fmaf(ex, ex, ey * ey));
                    float d = match[(size_t)i * n * m + (size_t)k * n + j] * (1.0f / sqrtf(fmaxf(s, 1e-20f)));
                    sx = fmaf(ex, d, sx);
                    sy = fmaf(ey, d, sy);
                    sz = fmaf(ez, d, sz);
                }
                sg[t * 3 + 0] = sx;
                sg[t * 3 + 1] = sy;
                sg[t * 3 + 2] = sz;
            }
            for (int j = 1; j < T; j <<= 1)
                for (int t = 0; t < T; ++t)
                    if ((t & j) == 0 && t + j < T && (t & (j - 1)) == 0) {
                        sg[t * 3 + 0] += sg[(t + j) * 3 + 0];
                        sg[t * 3 + 1] += sg[(t + j) * 3 + 1];
                        sg[t * 3 + 2] += sg[(t + j) * 3 + 2];
                    }
            grad2[((size_t)i * m + k) * 3 + 0] = sg[0];
            grad2[((size_t)i * m + k) * 3 + 1] = sg[1];
            grad2[((size_t)i * m + k) * 3 + 2] = sg[2];
        }
    }
}

/* ---------------------------------------------------------------------------------------
 * TargetNetwork.forward, model/target_network.py:31-38, for one sample:
 *   x[N,3] -> relu(x W1^T + b1) -> ... -> x W_L^T + b_L     (no activation on the output)
 * Flat weight layout per layer: W[out][in] row-major, then b[out] when use_bias
 * (model/target_network.py:40-45).  dims = {3, out_ch..., 3}, n_layers = len(dims)-1.
 * Accumulation: fp32, ascending input-channel order, bias added after the dot product
 * (torch.mm then '+ bias', target_network.py:33-36).
 * ------------------------------------------------------------------------------------- */
HP_EXPORT void hp_oracle_target_network_forward(int b, int npts, int n_layers, const int *dims,
                                                int use_bias, const float *weights,
                                                long long weight_stride, const float *points,
                                                float *out) {
#pragma omp parallel for schedule(static)
    for (int s = 0; s < b; ++s) {
        const float *w0 = weights + (size_t)s * weight_stride;
        float cur[512], nxt[512];
        for (int p = 0; p < npts; ++p) {
            const float *x = points + ((size_t)s * npts + p) * 3;
            cur[0] = x[0];
            cur[1] = x[1];
            cur[2] = x[2];
            const float *w = w0;
            for (int l = 0; l < n_layers; ++l) {
                int in = dims[l], on = dims[l + 1];
                const float *bias = w + (size_t)in * on;
                for (int o = 0; o < on; ++o) {
                    float acc = 0;
                    for (int c = 0; c < in; ++c) acc = fmaf(cur[c], w[(size_t)o * in + c], acc);
                    if (use_bias) acc += bias[o];
                    if (l + 1 < n_layers) acc = acc > 0 ? acc : 0;
                    nxt[o] = acc;
                }
                memcpy(cur, nxt, sizeof(float) * on);
                w += (size_t)in * on + (use_bias ? on : 0);
            }
            float *y = out + ((size_t)s * npts + p) * 3;
            y[0] = cur[0];
            y[1] = cur[1];
            y[2] = cur[2];
        }
    }
}

HP_EXPORT int hp_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
