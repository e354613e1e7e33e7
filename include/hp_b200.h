/*
 * hp_b200.h -- C ABI of the B200-native (sm_100a) point-set hot path of HyperPocket
 * (gmum/3d-point-clouds-autocomplete).  Shared library: libhp_b200.so.
 *
 * This is the drop-in boundary.  Everything is `extern "C"`, plain device
 * pointers and sizes, no torch types.  Conventions shared by every entry point:
 *   - all pointers are DEVICE pointers on the current CUDA device unless a name ends in
 *     `_host`; all tensors are dense, row-major, fp32 (indices int32);
 *   - `stream` is a `cudaStream_t` passed as `void*`; work is enqueued on it and the call
 *     returns without synchronising (fully stream-ordered, no legacy-stream memset --
 *     the reference's nndistancegrad() issues cudaMemset on the legacy stream,
 *     nndistance.cu:156-157; that hazard is not reproduced);
 *   - the library never allocates device memory (one documented exception: hp_nndistance, whose reference signature has no
 *     workspace argument): outputs and workspaces are caller-owned
 *     (the reference's glue allocates outputs with torch::empty, structural_loss.cpp:32-33,
 *     49,64-65,90-93,111-112; our Python glue does the same);
 *   - return value: HP_OK (0) or an HP_ERR_* code; the reference's launchers return void and
 *     throw std::runtime_error on launch failure (approxmatch.cu:334-337) or do not check at
 *     all (nndistance.cu:131-134).  hp_last_error_message() gives the detail string.
 *
 * The first five functions have exactly the argument lists of the reference's internal
 * launchers declared at utils/pytorch_structural_losses/structural_loss.cpp:11-15, which is
 * what a maintainer would bind.
 */
#ifndef HP_B200_H_
#define HP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HP_B200_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define HP_API __attribute__((visibility("default")))
#else
#define HP_API
#endif

enum {
    HP_OK = 0,
    HP_ERR_INVALID_ARGUMENT = 1, /* negative size, null pointer, misaligned workspace ... */
    HP_ERR_CUDA = 2,             /* a CUDA runtime call or kernel launch failed */
    HP_ERR_UNSUPPORTED = 3,      /* shape outside what the kernels support (documented per call) */
    HP_ERR_WORKSPACE = 4         /* workspace_bytes smaller than hp_*_workspace_bytes() */
};

HP_API int hp_version(void);
HP_API const char *hp_error_string(int code);
/* Thread-local detail of the last failing call on this thread ("" if none). */
HP_API const char *hp_last_error_message(void);

/* ------------------------------------------------------------------------------------
 * (a) Chamfer / nearest-neighbour distance
 * ---------------------------------------------------------------------------------- */

/* Replaces `void nndistance(int b,int n,const float*xyz,int m,const float*xyz2,float*result,
 * int*result_i,float*result2,int*result2_i,cudaStream_t)`  (structural_loss.cpp:14,
 * nndistance.cu:131-134).
 *   result [b,n]  = min_k |xyz[b,j]-xyz2[b,k]|^2,  result_i = argmin (lowest index on ties)
 *   result2[b,m], result2_i: the reverse direction.
 * d = fma(dz,dz,fma(dx,dx,dy*dy)) in fp32, bit-identical to the reference build.
 * One launch sequence covers both directions.  n==0 or m==0 with b>0 -> HP_ERR_INVALID_ARGUMENT
 * (the reference leaves the outputs uninitialised).
 * Input domain (all nearest-neighbour entry points): coordinates must be finite with |x| < 1e15 -- the ring kernels pad
 * partial tiles with points at +-1e18 and move running minima by one ulp for the tie rule; inf / NaN / larger magnitudes
 * give undefined distances and indices (not validated on the device).
 * The reference signature has no workspace argument: the call takes hp_chamfer_workspace_bytes() from the stream-ordered
 * pool of the current device (cudaMallocAsync / cudaFreeAsync on `stream`: no synchronisation, capturable) and runs the
 * same ring kernels as hp_nndistance_ws; it is the ONLY entry point of the library that allocates.  If the pool is
 * unavailable the workspace-free ordered-pair kernel is used (same bits, about half the speed). */
HP_API int hp_nndistance(int b, int n, const float *xyz, int m, const float *xyz2, float *result,
                  int *result_i, float *result2, int *result2_i, void *stream);

/* hp_nndistance with a caller-provided workspace (hp_chamfer_workspace_bytes(b,n,m) bytes, 16-byte aligned,
 * zero-filled once when allocated; the kernels restore that state).  Runs the "warp ring" kernels, which
 * evaluate every UNORDERED point pair once for both directions (d is bit-symmetric).  Identical results to hp_nndistance,
 * without its pool allocation and memset.  Coordinates must be finite with |x| < 1e15. */
HP_API int hp_nndistance_ws(int b, int n, const float *xyz, int m, const float *xyz2, float *result,
                     int *result_i, float *result2, int *result2_i, void *workspace, size_t workspace_bytes,
                     void *stream);

/* Replaces `void nndistancegrad(int b,int n,const float*xyz1,int m,const float*xyz2,
 * const float*grad_dist1,const int*idx1,const float*grad_dist2,const int*idx2,
 * float*grad_xyz1,float*grad_xyz2,cudaStream_t)`  (structural_loss.cpp:15,
 * nndistance.cu:155-160).
 *   grad_xyz1[b,j] = 2 g1[j] (a_j - b_idx1[j]) + sum_{k: idx2[k]==j} 2 g2[k] (a_j - b_k)
 *   grad_xyz2 symmetric.  Outputs are fully overwritten (no memset needed).
 * Atomic-free and deterministic (per-cloud stable counting sort of the index map in shared
 * memory, then a fixed-order gather) when n+m <= HP_NNGRAD_SMEM_POINTS; above that an
 * atomicAdd kernel like the reference's is used (non-deterministic summation order). */
HP_API int hp_nndistancegrad(int b, int n, const float *xyz1, int m, const float *xyz2,
                      const float *grad_dist1, const int *idx1, const float *grad_dist2,
                      const int *idx2, float *grad_xyz1, float *grad_xyz2, void *stream);
#define HP_NNGRAD_SMEM_POINTS 24576

/* Fused ChamferLoss forward (losses/champfer_loss.py:11-17 semantics, direct-form distances):
 * hp_nndistance plus loss[0] = sum(result) + sum(result2), reduced deterministically inside
 * the same launch.  `workspace` must hold hp_chamfer_workspace_bytes(b,n,m) bytes, be
 * 16-byte aligned and ZERO-FILLED once when allocated (the kernel restores that state, so it
 * can be reused across calls on one stream). */
HP_API size_t hp_chamfer_workspace_bytes(int b, int n, int m);
HP_API int hp_chamfer_forward(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1,
                       int *idx1, float *dist2, int *idx2, float *loss, void *workspace,
                       size_t workspace_bytes, void *stream);

/* Backward of the fused loss: like hp_nndistancegrad with grad_dist1 == grad_dist2 ==
 * grad_loss[0] (a single device scalar, read by the kernel -- no host sync). */
HP_API int hp_chamfer_backward(int b, int n, const float *xyz1, int m, const float *xyz2,
                        const int *idx1, const int *idx2, const float *grad_loss,
                        float *grad_xyz1, float *grad_xyz2, void *stream);

/* Replaces `ChamferLoss.batch_pairwise_dist(x, y)` (losses/champfer_loss.py:19-35), which callers such as
 * `dist_chamfer` (utils/metrics.py:78-83) use to get the whole matrix:
 *   P[b,i,j] = (|x_i|^2 + |y_j|^2) - 2 x_i.y_j   -- the reference's expansion form (may be slightly negative), fp32,
 *   x [b,nx,3], y [b,ny,3], P [b,nx,ny] (4*b*nx*ny bytes, written once; the kernel is HBM-write bound).
 * The reference builds it from three bmm and two discarded Gram matrices.  None of this library's hot paths needs P. */
HP_API int hp_batch_pairwise_dist(int b, int nx, int ny, const float *x, const float *y, float *P, void *stream);

/* Training-step pair: the forward additionally emits the INVERSE of both index maps (sorted in shared memory at the
 * tail of the forward's unpack kernel), so that the backward is a pure gather: one thread per point, no sort, no
 * atomics, same summation order (ascending source index) and bit-identical gradients to hp_chamfer_backward.
 *   inv1: hp_chamfer_inverse_ints(b,n,m,1) = b*(n+2m) ints, per cloud [perm1[n] | begin1[m] | end1[m]]: rows i
 *         sorted by (idx1[i], i), and for every column k its bucket perm1[begin1[k] .. end1[k]);
 *   inv2: hp_chamfer_inverse_ints(b,n,m,2) = b*(m+2n) ints, the same for columns sorted by (idx2[k], k).
 * hp_chamfer_inverse_ints returns 0 when the pair is unavailable (a cloud above 32768 points): use
 * hp_chamfer_forward / hp_chamfer_backward then.  Workspace as for hp_chamfer_forward. */
HP_API size_t hp_chamfer_inverse_ints(int b, int n, int m, int which);
HP_API int hp_chamfer_forward_inv(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1,
                           int *idx1, float *dist2, int *idx2, float *loss, int *inv1, int *inv2,
                           void *workspace, size_t workspace_bytes, void *stream);
HP_API int hp_chamfer_backward_inv(int b, int n, const float *xyz1, int m, const float *xyz2, const int *idx1,
                            const int *idx2, const int *inv1, const int *inv2, const float *grad_loss,
                            float *grad_xyz1, float *grad_xyz2, void *stream);
/* One training step of the fused loss in TWO kernels: loss = ChamferLoss(xyz1, xyz2) (losses/champfer_loss.py:11-17)
 * and its gradients for the upstream scalar grad_loss[0] (device memory, known before the forward is enqueued -- e.g. the
 * trainer's constant loss coefficient, core/epoch_loops.py:25-26).  The ring kernel is followed by a single tail kernel
 * (launched programmatically dependent; its CTAs wait for per-cloud tickets of the ring kernel instead of the whole grid, so
 * the tails of the early clouds run under the ring kernel's last wave) that decodes distances and indices, reduces the loss
 * in a fixed order, inverts both index maps in shared memory and gathers both gradients.
 * Outputs are bit-identical to hp_chamfer_forward_inv + hp_chamfer_backward_inv.  hp_chamfer_step_supported returns 0
 * for clouds above 4096 points (the tail kernel keeps a direction's keys in registers).  Workspace as for
 * hp_chamfer_forward.  Both kernels must be enqueued by this call on one stream (the tail relies on the ring kernel's CTAs
 * being resident or complete when it starts). */
HP_API int hp_chamfer_step_supported(int b, int n, int m);
HP_API int hp_chamfer_step(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_loss,
                    float *dist1, int *idx1, float *dist2, int *idx2, float *loss, float *grad_xyz1,
                    float *grad_xyz2, void *workspace, size_t workspace_bytes, void *stream);
/* hp_nndistancegrad (per-point upstream gradients, nn_distance.py:28-39) as the same gather over inverse maps from
 * hp_chamfer_forward_inv (whose `loss` may be NULL when only distances and indices are wanted). */
HP_API int hp_nndistancegrad_inv(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_dist1,
                          const int *idx1, const float *grad_dist2, const int *idx2, const int *inv1,
                          const int *inv2, float *grad_xyz1, float *grad_xyz2, void *stream);

/* ------------------------------------------------------------------------------------
 * (b) Approximate EMD (soft auction)
 * ---------------------------------------------------------------------------------- */

/* Replaces `void approxmatch(int b,int n,int m,const float*xyz1,const float*xyz2,float*match,
 * float*temp,cudaStream_t)` (structural_loss.cpp:11, approxmatch.cu:330-338).
 *   match [b,m,n] (match[b,l,k]: l indexes xyz2, k indexes xyz1), temp [b,2(n+m)] scratch
 *   (remainL,remainR,ratioL,ratioR per cloud; returned like the reference does). */
HP_API int hp_approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2, float *match,
                   float *temp, void *stream);

/* hp_approxmatch with a caller-provided workspace (hp_approxmatch_workspace_bytes, 16-byte aligned, contents
 * irrelevant): the auction records its per-level ratios there and `match` is written ONCE at the end (4nm bytes of
 * HBM traffic per cloud instead of nine read-modify-write sweeps); bit-identical results, about 8x faster at
 * B=32, 2048x2048. */
HP_API size_t hp_approxmatch_workspace_bytes(int b, int n, int m);
HP_API int hp_approxmatch_ws(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, float *temp,
                      void *workspace, size_t workspace_bytes, void *stream);

/* Replaces `void matchcost(...)` (structural_loss.cpp:12, approxmatch.cu:340-347):
 *   out[b] = sum_{l,k} match[b,l,k] * |xyz1[b,k]-xyz2[b,l]|. */
HP_API int hp_matchcost(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match,
                 float *out, void *stream);

/* Replaces `void matchcostgrad(...)` (structural_loss.cpp:13, approxmatch.cu:349-357):
 *   grad1[b,k] = sum_l match (x1-x2)/max(|x1-x2|,1e-10), grad2[b,l] the reverse. */
HP_API int hp_matchcostgrad(int b, int n, int m, const float *xyz1, const float *xyz2,
                     const float *match, float *grad1, float *grad2, void *stream);

/* Fused, match-free EMD cost for the metrics path (utils/metrics.py:71-76,147: match_cost is
 * only ever used forward-only there): cost[p] = match_cost(first[ia[p]], second[ib[p]]) for
 * `pairs` cloud pairs addressed through index lists (ia/ib may be NULL = identity).
 *   first [na, npts, 3], second [nb, npts, 3].  workspace: hp_emd_cost_workspace_bytes().
 * Points of the second cloud whose mass is exhausted (remainR == 0 after the clamp of approxmatch.cu:140) are left
 * out of every later pass: their terms are exact zeros, so no sum changes (csrc/emd.cu: emd_compact_kernel); the
 * running time therefore depends on the data, the result does not. */
HP_API size_t hp_emd_cost_workspace_bytes(int pairs, int n, int m);
HP_API int hp_emd_cost_pairs(int pairs, int n, int m, const float *first, const int *ia,
                      const float *second, const int *ib, float *cost, void *workspace,
                      size_t workspace_bytes, void *stream);
/* Opt-in shortcut, NOT within the 1e-5 parity bar: the third pass of a level and the first pass of the next share one
 * ex2 (e = e'^4): 3 instead of 4 MUFU operations per point pair and level, 19 launches instead of 27.  It was
 * 1.2x faster than hp_emd_cost_pairs until that learned to leave exhausted points out; the shortcut sweeps every
 * point and is now the SLOWER of the two (1.25 vs 1.18 ms at B=32, 2048^2): kept as a measured variant only.
 * Measured worst deviation of the cost from the reference extension: 2.1e-5 relative (hp_emd_cost_pairs: 1.4e-6), see
 * csrc/emd.cu.  Same arguments and workspace as hp_emd_cost_pairs. */
HP_API int hp_emd_cost_pairs_fast(int pairs, int n, int m, const float *first, const int *ia,
                           const float *second, const int *ib, float *cost, void *workspace,
                           size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------
 * (c) TargetNetwork: the per-sample MLP whose weights the hypernetwork emits
 *     (model/target_network.py:5-45, driven by the loop at model/full_model.py:67-74).
 * ---------------------------------------------------------------------------------- */
/* Layer widths: dims_host[0..n_layers] (HOST array) = {3, layer_out_channels..., 3}.  Per sample the flat
 * weight vector holds, for each layer, W[out][in] row-major then (use_bias) b[out]
 * (target_network.py:40-45); hp_target_network_num_weights() is its length (19011 for 32,64,128,64). */
HP_API long long hp_target_network_num_weights(int n_layers, const int *dims_host, int use_bias);

/* Arithmetic of the tuned 3,32,64,128,64,3 kernels (process-wide; other widths always run the generic FP32 kernels):
 *   0 (default)  error-compensated 3xTF32 on the tensor cores (every operand split into hi + lo, three tensor-core products per fp32
 *                product, fp32 accumulation; weight gradients accumulated per 128-point tile and summed in fp32): within 3e-6 of the
 *                fp32 torch.mm chain of target_network.py:31-38 (measured, profiles/r02_tn_error_margins.txt; bar 1e-5).  The forward
 *                runs on tcgen05 (kind::tf32, accumulators and activations in tensor memory: it allocates all 512 columns of the
 *                SM's tensor memory while it runs), the backward on mma.sync with the running gradient in tensor memory;
 *   1            plain fp32 FFMA chains on the CUDA cores (round-1 kernels);
 *   2            as 0 with the forward on mma.sync as well (no tcgen05.mma).
 * Returns HP_ERR_INVALID_ARGUMENT for any other value. */
HP_API int hp_target_network_set_mode(int mode);

/* out[s] = TargetNetwork(weights[s]).forward(points[s]) for s in [0,b): ReLU after every layer but the
 * last (target_network.py:31-38).  ONE launch for the whole batch, no activation goes to HBM.
 *   weights [b, W];  points [b, n, 3] with points_batch_stride = 3n floats, or one shared cloud
 *   [n, 3] with points_batch_stride = 0;  out [b, n, 3] (channels_first = 0) or [b, 3, n]
 *   (channels_first = 1: the layout FullModel.forward writes, full_model.py:68,74).
 * Widths 3,32,64,128,64,3 run the tuned kernel; any other widths a generic one (HP_ERR_UNSUPPORTED if a
 * layer is too wide for shared memory). */
HP_API int hp_target_network_forward(int b, int n, int n_layers, const int *dims_host, int use_bias,
                              const float *weights, const float *points, long long points_batch_stride,
                              float *out, int channels_first, void *stream);

/* Gradient of sum(out * grad_out) w.r.t. the flat weights (what autograd hands the hypernetwork) and,
 * when grad_points != NULL (needs per-sample points), w.r.t. the input points.  The forward is recomputed
 * tile by tile inside the kernel; nothing was stashed.  Deterministic (fixed-order reductions, no float
 * atomics).  grad_out has the layout of `out` (channels_first).  grad_weights [b, W] is fully overwritten.
 * `workspace`: hp_target_network_backward_workspace_bytes() bytes, 16-byte aligned, contents irrelevant
 * (per-CTA partial gradients and arrival counters; the call initialises what it needs on `stream`). */
HP_API size_t hp_target_network_backward_workspace_bytes(int b, int n, int n_layers, const int *dims_host, int use_bias);
HP_API int hp_target_network_backward(int b, int n, int n_layers, const int *dims_host, int use_bias,
                               const float *weights, const float *points, long long points_batch_stride,
                               const float *grad_out, int channels_first, float *grad_weights,
                               float *grad_points, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------
 * (d) Set-vs-set metrics (utils/metrics.py:121-158): one row block of the cloud-distance
 *     matrix per call; rows are what gets sharded across GPUs.
 * ---------------------------------------------------------------------------------- */

/* cd[r - row_begin, s] = mean_i min_j d(first[r,i], second[s,j]) + mean_j min_i d(...)
 * for r in [row_begin,row_end), s in [0,nb).  first [na,n,3], second [nb,m,3].
 * Each unordered point pair is evaluated once and feeds both minima. */
HP_API int hp_pairwise_cd(int na, int nb, int n, int m, const float *first, const float *second,
                   int row_begin, int row_end, float *cd, void *stream);

/* The same cloud distance for an explicit list of cloud pairs: cd[p] = CD(first[pair_r[p]], second[pair_s[p]]),
 * pair_r / pair_s device int32 [npairs] (indices are not range-checked).  Used for the SYMMETRIC matrices of the
 * 1-NN two-sample test (knn, utils/metrics.py:162-191: M_xx and M_yy compare a set with itself), where only the
 * strict upper triangle is evaluated and the pair list is what gets sharded across GPUs. */
HP_API int hp_pairwise_cd_pairs(long long npairs, int n, int m, const float *first, const float *second,
                         const int *pair_r, const int *pair_s, float *cd, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HP_B200_H_ */
