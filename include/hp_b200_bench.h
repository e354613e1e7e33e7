/*
 * hp_b200_bench.h -- measurement helpers of libhp_b200_bench.so (NOT part of the product library libhp_b200.so).
 *
 * libhp_b200_bench.so is the product's sources compiled once more with -DHP_BENCH_BUILD: it exports everything
 * hp_b200.h declares plus the functions below, and honours the A/B environment switches (HP_RING_VARIANT, HP_NO_PDL,
 * HP_NN_RING, HP_NN_VARIANT) that the product library does not contain.  bench.py and tools/ load it for the roofline
 * denominators and for timing the dominant kernel alone; nothing under the package imports it.
 */
#ifndef HP_B200_BENCH_H_
#define HP_B200_BENCH_H_

#include "hp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Runs a register-resident FFMA (kind 0), packed FFMA2 (kind 1) or MUFU.EX2 (kind 2) chain, the Chamfer inner-loop
 * instruction mix (kinds 3-5), legacy mma.sync TF32 (kind 6), or a register-only replica of the ring kernel's rotation
 * (kinds 7-12: FMA-pipe ops only / + FMNMX3 / + FSETP,SEL bookkeeping / the latter with 3, 2, 1 warps per scheduler; see
 * csrc/api.cu and DESIGN.md 4.1) on every SM and returns the achieved rate in *rate (FLOP/s for kinds 0-1 and 3-6, ex2/s
 * for kind 2, packed instructions per lane per second for kinds 7-12), timed with CUDA events on `stream` (synchronises). */
HP_API int hp_measure_peak(int kind, int iters, double *rate_host, void *stream);
/* Launches ONLY the dominant kernel of the Chamfer step (nn_ring_kernel: all-pairs distances, both directions, keys into the
 * workspace) so that bench.py can time it alone with CUDA events.  The keys are left in `workspace` (zero-filled on entry,
 * hp_chamfer_workspace_bytes): use a private workspace and discard it afterwards. */
HP_API int hp_measure_chamfer_ring_only(int b, int n, const float *xyz1, int m, const float *xyz2, void *workspace,
                                 size_t workspace_bytes, void *stream);

/* Per-CTA timeline of the Chamfer step: while a device buffer is set (NULL switches it off), every nn_ring_kernel CTA writes
 * its %globaltimer at start / end to slots [2*cta, 2*cta+1] and every nn_ring_tail_kernel CTA its %globaltimer
 * at ten points of its life to slots [2*ring_ctas + 10*cta ...] (nanoseconds; see the HP_TRACE marks in csrc/chamfer_ring.cu).  tools/chamfer_timeline.py prints the summary. */
HP_API int hp_measure_set_trace(void *device_u64_buffer);

#ifdef __cplusplus
}
#endif
#endif /* HP_B200_BENCH_H_ */
