#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native HyperPocket point-set hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], named in `config.workload`): Chamfer nearest-neighbour
distance forward + backward, B=32, N=M=2048, fp32, synthetic clouds ~ U[-0.5,0.5]^3.
One "step" = Chamfer forward (both directions + loss) + backward (both gradients), captured once as a CUDA
graph (ChamferStepGraph) and replayed: ring kernel + one fused tail kernel (loss, inverse maps, both gradients).
metric = ordered (query, candidate) point pairs evaluated per second = 2*B*N*M / t_step.

One JSON line on stdout (rank 0).  Keys follow the driver contract; in addition
  roofline     : the dominant kernel against the FP32 FFMA peak MEASURED LIVE in this run
                 (MEASURED_PEAKS.json carries no FP32 number; K=3 keeps the path off the tensor cores); carries
                 `metrics_eval` (config C5: the sharded evaluation at this N) so that the driver keeps it
  cpu_baseline : the reference's own pure-torch ChamferLoss (unmodified, from baseline/_ref; the oracle port when that is not
                 staged) on this box's host cores; `cpu_baseline_c1` = BASELINE config C1 (reference FullModel forward +
                 ChamferLoss forward + backward on CPU)
  e2e          : the same step through the public Python API with pinned HOST buffers, DEPENDENT steps (one after the other, as
                 a trainer runs them): H2D of both clouds, the step, D2H of the loss -- the gradients stay on the device for
                 the TargetNetwork backward.  The pipelined rate over independent steps and the variant that also copies both
                 gradients back are reported beside it.
  secondary_comparators : the reference's own CUDA extension (oracle/_ref, unmodified) and its pure-torch ChamferLoss timed on
                 this GPU at C2 / C3; c4_full_step: the whole training step (encoder + hypernetwork + fused TargetNetwork +
                 Chamfer fwd/bwd + backward + Adam) as one CUDA graph next to the reference's eager step on this GPU.
N > 1: Chamfer does not shard ("replicas only") -> every rank runs an independent
replica of the workload ("scaling": "weak"), no data-path collective.  The part of BASELINE.json's metric that
DOES shard -- the all-pairs MMD/COV/1-NNA evaluation (config C5) -- is measured at the same N and reported in
`metrics_eval` (strong scaling, rows of the cloud-distance matrices sharded over the ranks, NCCL gathers of the
per-row / per-column minima only).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

B, N, M = 32, 2048, 2048
PAIRS_PER_STEP = 2 * B * N * M  # ordered (query, candidate) evaluations, both directions
FLOP_PER_PAIR = 8  # 3 sub, 3 mul, 2 add
NCU_RING_DRAM_BYTES = 2658816  # profiles/r0*_ncu_full_nn_ring.txt (dram__bytes_read.sum + dram__bytes_write.sum, per launch)
METRIC = "chamfer_point_pairs_per_s"
UNIT = "pairs/s"
WORKLOAD = f"chamfer_nn_distance_fwd+bwd_B{B}_N{N}_M{M}_fp32"


def _config():
    """Identical in both arms (the driver compares the dicts): names the workload, nothing run-specific."""
    return {"workload": WORKLOAD, "batch": B, "points_a": N, "points_b": M,
            "l2_policy": "GPU arm: L2 flushed (256 MiB write) between timed iterations; CPU arm: n/a",
            "parallelism": "replicas (no collective on this path); the sharded C5 evaluation is in roofline.metrics_eval"}


class ClockSampler:
    """Samples SM clocks / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self._nv = None
            self.error = repr(e)

    def _run(self):
        nv = self._nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def _synthetic(torch, seed):
    g = torch.Generator().manual_seed(seed)
    a = (torch.rand(B, N, 3, generator=g) - 0.5).contiguous()
    b = (torch.rand(B, M, 3, generator=g) - 0.5).contiguous()
    return a, b


REF_SAMPLE_CLOUDS = 16  # clouds of the 32-cloud batch per reference-arm step (bounded sample; the CPU cost is linear in clouds)


def _reference_chamfer():
    """(callable loss(preds, gts), kind): the reference's own ChamferLoss module when baseline/_ref is staged, else the port."""
    try:
        from baseline import ref_loader

        if ref_loader.ref_root() is not None:
            mod = ref_loader.reference_chamfer_loss()()
            mod.use_cuda = False
            return mod, "reference"
    except Exception:
        pass
    from oracle import oracle as O

    return O.chamfer_loss_torch, "port"


def cpu_reference_arm(steps: int, warmup: int, clouds: int = B):
    """The reference's own CPU implementation of the path: pure-torch ChamferLoss (losses/champfer_loss.py:11-35: three bmm,
    two min, sum) forward + backward on all host threads, on `clouds` of the workload's B clouds per step (cycling through
    the batch).  Returns (seconds per step list, threads, kind)."""
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    loss_fn, kind = _reference_chamfer()
    a_all, b_all = _synthetic(torch, 0)

    def step(i):
        lo = (i * clouds) % B
        a = a_all[lo:lo + clouds].clone().requires_grad_(True)
        b = b_all[lo:lo + clouds].clone().requires_grad_(True)
        loss = loss_fn(b, a)  # forward(preds, gts)
        loss.backward()
        return float(loss.detach())

    for i in range(max(0, warmup)):
        step(i)
    times = []
    for i in range(steps):
        t0 = time.perf_counter()
        step(i)
        times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads(), kind


def cpu_reference_c1(reps: int = 2):
    """BASELINE config C1 / BASELINE.md 4: the reference's own FullModel (HyperPocket mode, 3D-EPN airplane settings) forward +
    pure-torch ChamferLoss forward + full backward on CPU, B=32, 1024-point partial clouds -> 2048-point completion."""
    import json as _json

    import torch

    from baseline import ref_loader

    if ref_loader.ref_root() is None:
        return None
    tree = ref_loader.reference_tree()
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(1856)
    model = tree.RefFullModel(_json.loads(_json.dumps(ref_loader.full_model_config("config_3depn_airplane.json.sample"))))
    model.apply(ref_loader.weights_init)
    model.train()
    loss_fn = tree.RefChamferLoss()
    loss_fn.use_cuda = False
    g = torch.Generator().manual_seed(0)
    existing, missing = torch.rand(32, 1024, 3, generator=g) - 0.5, torch.rand(32, 1024, 3, generator=g) - 0.5
    gt = torch.rand(32, 2048, 3, generator=g) - 0.5
    times = []
    for rep in range(reps + 1):
        for p in model.parameters():
            p.grad = None
        t0 = time.perf_counter()
        rec, logvar, mu = model(existing.clone(), missing.clone(), list(gt.shape), 1, torch.device("cpu"))
        t1 = time.perf_counter()
        loss = torch.mean(0.05 * loss_fn(gt, rec.permute(0, 2, 1)))
        t2 = time.perf_counter()
        loss.backward()
        t3 = time.perf_counter()
        if rep > 0:
            times.append((t1 - t0, t2 - t1, t3 - t2))
    fwd, cham, bwd = (statistics.mean(t[i] for t in times) for i in range(3))
    return {"what": "BASELINE config C1: reference FullModel.forward (HyperPocket, B=32, 1024 -> 2048 points) + reference pure-torch "
                    "ChamferLoss forward + loss.backward() through everything, CPU, unmodified files from baseline/_ref",
            "full_model_forward_s": fwd, "chamfer_forward_s": cham, "backward_s": bwd, "step_s": fwd + cham + bwd,
            "cores": torch.get_num_threads(), "reps": len(times), "kind": "reference",
            "chamfer_unordered_pairs_per_s_fwd": 32 * 2048 * 2048 / cham}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    times, cores, kind = cpu_reference_arm(steps, warmup, REF_SAMPLE_CLOUDS)
    t = statistics.mean(times)
    pairs = 2 * REF_SAMPLE_CLOUDS * N * M
    value = pairs / t
    sample = (f"each step = {REF_SAMPLE_CLOUDS} of the workload's {B} clouds (N=M={N}), cycling through the batch: "
              "the reference's pure-torch ChamferLoss forward + backward (expansion form, 3 bmm + 2 min) on CPU; "
              "the cost is linear in the number of clouds, so pairs/s equals the full-batch rate")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "pairs_per_step": pairs, "full_batch_ms_per_step_equivalent": t * 1e3 * B / REF_SAMPLE_CLOUDS},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU arm; " + ("unmodified reference module from baseline/_ref" if kind == "reference" else "oracle port (baseline/_ref not staged)"),
    }
    print(json.dumps(line), flush=True)


def _events_timed(torch, fn, steps, warmup, flush, stream, barrier):
    """CUDA-event time of each of `steps` calls of fn on `stream`; L2 evicted (outside the timed interval) before each."""
    for _ in range(warmup):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    for e0, e1 in evs:
        flush.fill_(1)  # evict L2 (126 MB); the CPU enqueues the timed work while this runs, so no launch gap is timed
        e0.record(stream)
        fn()
        e1.record(stream)
    barrier()
    return [e0.elapsed_time(e1) for e0, e1 in evs]  # ms


def _other_paths(torch, hp, dev, fp32_peak, mufu_peak, hbm_peak_gbs, flush, stream, barrier):
    """Secondary hot-path rows of SURVEY 8 (TargetNetwork C4, EMD C3, pairwise CD C5 sample, a2 matrix, f4 head): time + roofline."""
    out = {}
    LOC = [32, 64, 128, 64]
    tb, tn = 64, 2048
    tng = hp.TargetNetworkStepGraph(tb, tn, LOC, True, dev, channels_first=True)
    g = torch.Generator().manual_seed(5)
    tng.weights.copy_((torch.randn(tb, 19011, generator=g) * 0.15).to(dev))
    ms = statistics.mean(_events_timed(torch, tng.replay, 20, 3, flush, stream, barrier))
    flop = (37440.0 + 74688.0) * tb * tn  # algorithmic: fwd + bwd (the in-kernel forward recompute is not counted)
    executed = flop + 37440.0 * tb * tn
    tf32_peak = hp._native.measure_peak(6, 4096)  # legacy mma.sync tf32 FLOP/s, measured live (tcgen05 has no fp32-accurate input type)
    out["target_network_fwd+bwd_B64_N2048"] = {
        "ms": ms, "algorithmic_tflops": flop / ms / 1e9, "frac_fp32_peak": flop / (ms * 1e-3) / fp32_peak,
        "executed_frac_fp32_peak": executed / (ms * 1e-3) / fp32_peak,
        "arithmetic": "error-compensated 3xTF32 (three tensor-core products per fp32 product, fp32 accumulation): forward on tcgen05 with the "
                      "activations in tensor memory, backward on mma.sync with the running weight gradient in tensor memory: within 3e-6 of the fp32 chain, bar 1e-5 (tests, tools/tn_error_margins.py)",
        "mma_tf32_peak_tflops_measured": tf32_peak / 1e12,
        "frac_of_3xtf32_tensor_bound": 3.0 * executed / (ms * 1e-3) / tf32_peak,
        "what": "frac_fp32_peak is north_star's yardstick (algorithmic FLOP over the FP32 FFMA peak); frac_of_3xtf32_tensor_bound = "
                "executed FLOP x 3 products over the measured LEGACY mma.sync tf32 peak (the backward's pipe; the tcgen05 forward has 4x that)"}
    hp.target_network_set_mode("fp32")
    try:
        tng1 = hp.TargetNetworkStepGraph(tb, tn, LOC, True, dev, channels_first=True)
        tng1.weights.copy_(tng.weights)
        ms1 = statistics.mean(_events_timed(torch, tng1.replay, 20, 3, flush, stream, barrier))
        out["target_network_fwd+bwd_B64_N2048"]["fp32_ffma_mode_ms"] = ms1
        out["target_network_fwd+bwd_B64_N2048"]["fp32_ffma_mode_frac_fp32_peak"] = flop / (ms1 * 1e-3) / fp32_peak
        del tng1
    finally:
        hp.target_network_set_mode("tf32x3")
    hpg = hp.HotPathStepGraph(tb, tn, LOC, True, dev)
    hpg.weights.copy_(tng.weights)
    ms = statistics.mean(_events_timed(torch, hpg.replay, 20, 3, flush, stream, barrier))
    out["c4_hot_path_step_B64_N2048"] = {
        "ms": ms, "what": "fused TargetNetwork fwd -> Chamfer ring kernel -> Chamfer tail (loss + both gradients) -> TargetNetwork bwd, one CUDA graph "
                          "(config C4 without encoder / hypernetwork; the whole step is c4_full_step)",
        "algorithmic_tflops": ((37440.0 + 74688.0) * tb * tn + 16.0 * tb * tn * tn) / ms / 1e9}
    del tng, hpg
    torch.cuda.empty_cache()
    eb = 32
    a = (torch.rand(eb, 2048, 3, generator=g) - 0.5).to(dev)
    b = (torch.rand(eb, 2048, 3, generator=g) - 0.5).to(dev)
    ms = statistics.mean(_events_timed(torch, lambda: hp.emd_cost_pairs(a, b), 5, 2, flush, stream, barrier))
    ex2 = 27.0 * eb * 2048 * 2048
    out["emd_match_cost_fused_B32_2048x2048"] = {
        "ms": ms, "algorithmic_tex2_per_s": ex2 / ms / 1e9, "frac_mufu_peak": ex2 / (ms * 1e-3) / mufu_peak,
        "note": "ALGORITHMIC ex2 (27 per point pair, SURVEY 8d) over the measured MUFU.EX2 peak.  The kernels skip, exactly, the points of "
                "the second cloud whose mass is exhausted (zero-weight terms: no bit of any sum changes), so fewer ex2 are executed "
                "than counted and this fraction can exceed 1 on large calls"}
    ms = statistics.mean(_events_timed(torch, lambda: hp.match_cost(a, b), 3, 1, flush, stream, barrier))
    out["emd_approx_match+match_cost_B32_2048x2048"] = {"ms": ms}
    nr = 128
    ref = (torch.rand(nr, 2048, 3, generator=g) - 0.5).to(dev)
    smp = (torch.rand(nr, 2048, 3, generator=g) - 0.5).to(dev)
    ms = statistics.mean(_events_timed(torch, lambda: hp.pairwise_cd(ref, smp), 3, 1, flush, stream, barrier))
    pairs = float(nr) * nr * 2048 * 2048
    out["pairwise_cd_128x128_clouds_2048pts"] = {
        "ms": ms, "unordered_pairs_per_s": pairs / (ms * 1e-3),
        "frac_fp32_peak_8flop_per_unordered_pair": 8 * pairs / (ms * 1e-3) / fp32_peak,
        "frac_fp32_peak_16flop_per_unordered_pair": 16 * pairs / (ms * 1e-3) / fp32_peak,
        "note": "SURVEY 8d counts C5 as 8 FLOP per unordered pair (one evaluation feeds both minima): that is the roofline figure; "
                "16 FLOP per unordered pair (= 8 per ORDERED pair, the Chamfer convention) is printed beside it"}
    # a2: ChamferLoss.batch_pairwise_dist, the [B,N,M] matrix in one kernel (HBM-write bound)
    ms = statistics.mean(_events_timed(torch, lambda: hp.batch_pairwise_dist(a, b), 5, 2, flush, stream, barrier))
    nbytes = 4.0 * eb * 2048 * 2048
    out["batch_pairwise_dist_B32_2048x2048"] = {"ms": ms, "gbs_written": nbytes / (ms * 1e-3) / 1e9,
                                                "frac_hbm_peak": (nbytes / (ms * 1e-3) / 1e9 / hbm_peak_gbs) if hbm_peak_gbs else None}
    return out


def _hypernet_and_c4(torch, hp, dev, flush, stream, barrier):
    """f4 (fused hypernetwork head) and BASELINE config C4 as named: the WHOLE training step -- stock encoder + hypernetwork,
    fused TargetNetwork, Chamfer fwd/bwd, backward through everything, Adam -- captured as one CUDA graph, next to the
    reference's own eager step (its FullModel, its pure-torch ChamferLoss, its per-sample loop) on the same GPU."""
    import json as _json

    from baseline import ref_loader

    if ref_loader.ref_root() is None:
        return {"unavailable": "baseline/_ref not staged"}
    tree = ref_loader.reference_tree()
    cfg = ref_loader.full_model_config("config_completion.json.sample")  # Completion3D shape: HyperRec mode, batch 64 x 2048
    bsz, npts = 64, 2048
    out = {"what": "Completion3D-shape training step (settings/config_completion.json.sample: HyperRec mode), batch 64, "
                   "existing 2048 points, gt 2048 points, Adam(lr=1e-4), loss_coef 0.05"}
    g = torch.Generator().manual_seed(11)
    existing = (torch.rand(bsz, npts, 3, generator=g) - 0.5).to(dev)
    gt = (torch.rand(bsz, npts, 3, generator=g) - 0.5).to(dev)

    def make(cls):
        torch.manual_seed(1856)
        m = cls(_json.loads(_json.dumps(cfg)))
        m.apply(ref_loader.weights_init)
        return m.to(dev).train()

    # ---- reference eager step on this GPU (core/epoch_loops.py:14-39) ----
    ref_model = make(tree.RefFullModel)
    ref_opt = torch.optim.Adam(ref_model.parameters(), lr=1e-4)
    ref_loss = tree.RefChamferLoss().to(dev)

    def ref_step():
        ref_opt.zero_grad()
        rec, _lv, _mu = ref_model(existing.clone(), None, list(gt.shape), 1, dev)
        loss = torch.mean(0.05 * ref_loss(gt, rec.permute(0, 2, 1)))
        loss.backward()
        ref_opt.step()
        return loss

    ref_step()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(3):
        ref_step()
    torch.cuda.synchronize(dev)
    out["reference_eager_step_ms"] = (time.perf_counter() - t0) / 3 * 1e3
    del ref_opt, ref_loss
    # ---- f4: the head alone, reference 5 x Linear + cat vs one GEMM ----
    hn_ref = ref_model.hyper_network
    trunk_in = torch.randn(bsz, 128, device=dev)
    gout = torch.randn(bsz, 19011, device=dev)

    def head_fwd_bwd(hn):
        def f():
            for p in hn.parameters():
                p.grad = None
            y = hn(trunk_in)
            y.backward(gout)
        return f

    ms_ref = statistics.mean(_events_timed(torch, head_fwd_bwd(hn_ref), 10, 3, flush, stream, barrier))
    import copy as _copy

    hn_fused = hp.fuse_hypernetwork_head(_copy.deepcopy(hn_ref))
    ms_fused = statistics.mean(_events_timed(torch, head_fwd_bwd(hn_fused), 10, 3, flush, stream, barrier))
    err = float((hn_fused(trunk_in) - hn_ref(trunk_in)).abs().max() / hn_ref(trunk_in).abs().max())
    out["hypernetwork_fwd+bwd_B64"] = {"reference_5_linear_cat_ms": ms_ref, "fused_head_ms": ms_fused, "max_rel_diff": err,
                                       "what": "whole HyperNetwork (trunk + head) forward + backward, eager, fp32 cuBLAS"}
    del hn_fused, ref_model
    torch.cuda.empty_cache()
    # ---- ours: the whole step as one graph ----
    model = make(tree.OurFullModel)
    hp.fuse_hypernetwork_head(model.hyper_network)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True)
    step = hp.FullModelStepGraph(model, opt, bsz, npts, 0, npts, dev, loss_coef=0.05)
    step.existing.copy_(existing.transpose(1, 2))
    step.gt.copy_(gt)
    torch.manual_seed(5)
    step.load_points(1)
    ms = statistics.mean(_events_timed(torch, step.replay, 20, 3, flush, stream, barrier))
    out["ours_graph_step_ms"] = ms
    out["ours_graph_loss_r_finite"] = bool(torch.isfinite(step.loss_r).item())
    out["speedup_vs_reference_eager_same_gpu"] = out["reference_eager_step_ms"] / ms
    out["ours_what"] = ("FullModelStepGraph: stock Encoder + HyperNetwork trunk (cuDNN/cuBLAS), fused head GEMM, fused TargetNetwork fwd, "
                        "nn_ring_kernel + nn_ring_tail_kernel, fused TargetNetwork bwd, autograd through hypernetwork + encoder, "
                        "capturable Adam -- one graph replay per step; the host draws the TargetNetwork input points per step "
                        "(load_points, not in the timed replay: 64 x 2048 x 3 floats, one pinned H2D copy)")
    return out


def _secondary_comparators(torch, hp, dev, flush, stream, barrier):
    """the reference's own CUDA extension (unmodified, oracle/_ref) and its pure-torch ChamferLoss on THIS GPU."""
    out = {}
    try:
        from oracle import oracle as O

        ext = O.load_reference_ext()
    except Exception as e:  # pragma: no cover
        ext = None
        out["error"] = repr(e)
    g = torch.Generator().manual_seed(3)
    a = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev)
    b = (torch.rand(B, M, 3, generator=g) - 0.5).to(dev)
    if ext is not None:
        g1, g2 = torch.ones(B, N, device=dev), torch.ones(B, M, device=dev)

        def nn_fwd_bwd():
            d1, i1, d2, i2 = ext.NNDistance(a, b)
            ext.NNDistanceGrad(a, b, i1, i2, g1, g2)

        ms = statistics.mean(_events_timed(torch, nn_fwd_bwd, 10, 3, flush, stream, barrier))
        out["reference_ext_NNDistance+NNDistanceGrad_C2"] = {"ms": ms, "pairs_per_s": PAIRS_PER_STEP / (ms * 1e-3),
                                                            "what": "nndistance.cu:131-160 built for sm_100 from the reference's sources"}

        def emd():
            match, _t = ext.ApproxMatch(a, b)
            ext.MatchCost(a, b, match)

        ms = statistics.mean(_events_timed(torch, emd, 3, 1, flush, stream, barrier))
        out["reference_ext_ApproxMatch+MatchCost_C3"] = {"ms": ms, "what": "approxmatch.cu:330-347"}
    else:
        out["reference_ext"] = "oracle/_ref not built"
    try:
        from baseline import ref_loader

        if ref_loader.ref_root() is not None:
            mod = ref_loader.reference_chamfer_loss()().to(dev)
            ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)

            def torch_chamfer():
                ar.grad = br.grad = None
                mod(br, ar).backward()

            ms = statistics.mean(_events_timed(torch, torch_chamfer, 5, 2, flush, stream, barrier))
            out["reference_pure_torch_ChamferLoss_on_gpu_C2"] = {"ms": ms, "pairs_per_s": PAIRS_PER_STEP / (ms * 1e-3),
                                                                "what": "losses/champfer_loss.py:11-35 (3 bmm + 2 min) forward + backward, eager"}
            del ar, br
            torch.cuda.empty_cache()
    except Exception as e:  # pragma: no cover
        out["reference_pure_torch_error"] = repr(e)
    return out


def _metrics_eval(torch, dist, hp, dev, world, rank, barrier, mufu_peak, fp32_peak, emd_1nn):
    """Second half of BASELINE.json's metric: all-pairs MMD/COV/1-NNA evaluation, 1000 generated vs 1000 reference
    clouds x 2048 points (config C5), sharded over the ranks (strong scaling).  CD: full size with 1-NNA (the 1000x1000
    ref-vs-sample matrix + the upper triangles of the two self-distance matrices).  EMD: the full 1000x1000 ref-vs-sample matrix
    (MMD / COV) by default; with --metrics-emd-1nn also the two 1000x1000 self matrices of 1-NNA-EMD (3x the time)."""
    g = torch.Generator().manual_seed(1234)  # identical inputs on every rank
    smp = (torch.rand(1000, 2048, 3, generator=g) - 0.5).to(dev)
    ref = (torch.rand(1000, 2048, 3, generator=g) - 0.5).to(dev)
    out = {"clouds": "1000 generated vs 1000 reference x 2048 pts", "scaling": "strong", "n_gpus": world}

    def timed(fn):
        barrier()
        t0 = time.perf_counter()
        r = fn()
        {k: float(v) for k, v in r.items()}  # values are read on the host like core/experiments.py:97 does
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), r

    hp.compute_all_metrics(smp[:64], ref[:64], with_emd=False, one_nn=True)  # warm-up
    t, r = timed(lambda: hp.compute_all_metrics(smp, ref, with_emd=False, one_nn=True))
    out["cd_mmd_cov_1nna_s"] = t
    # cloud pairs actually evaluated: the full ref-vs-sample matrix + the strict upper triangles of the two symmetric
    # self-distance matrices (the reference formulation, three full matrices, would be 3 * 10^6)
    cloud_pairs = 1000 * 1000 + 2 * (1000 * 999 // 2)
    out["cd_cloud_pairs_evaluated"] = cloud_pairs
    out["cd_unordered_point_pairs_per_s"] = cloud_pairs * 2048.0 * 2048 / t
    out["cd_frac_fp32_peak_8flop_per_unordered_pair_all_gpus"] = 8 * cloud_pairs * 2048.0 * 2048 / t / (fp32_peak * world)
    out["1-NN-CD-acc"] = float(r["1-NN-CD-acc"])
    out["cov(Coverage)-CD"] = float(r["cov(Coverage)-CD"])
    hp.compute_all_metrics(smp[:16], ref[:16], with_emd=True, one_nn=False)
    t, r = timed(lambda: hp.compute_all_metrics(smp, ref, with_emd=True, one_nn=False))
    out["cd+emd_mmd_cov_1000x1000_s"] = t
    out["emd_cloud_pairs_per_s"] = 1000 * 1000 / t
    out["emd_frac_mufu_peak_27ex2_per_pair_all_gpus"] = 27.0 * 1e6 * 2048 * 2048 / t / (mufu_peak * world)
    out["emd_frac_note"] = ("algorithmic ex2 count over the MUFU peak; exhausted points are skipped exactly (DESIGN.md 4.4), so the executed "
                            "count is lower and the fraction may exceed 1")
    out["mmd(Fidelity)-EMD"] = float(r["mmd(Fidelity)-EMD"])
    out["cov(Coverage)-EMD"] = float(r["cov(Coverage)-EMD"])
    if emd_1nn:
        t, r = timed(lambda: hp.compute_all_metrics(smp, ref, with_emd=True, one_nn=True))
        out["cd+emd_mmd_cov_1nna_1000x1000_s"] = t
        out["1-NN-EMD-acc"] = float(r["1-NN-EMD-acc"])
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
    native = hp._native
    native.load()
    loss_mod = hp.ChamferLoss()

    a_h, b_h = _synthetic(torch, rank)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # The step, captured once as a CUDA graph (the repo's public API for launch-bound steps, graphs.py): the ring kernel
    # (all-pairs distances, both directions) + ONE tail kernel (distances/indices, loss, inverse maps, both gradients).
    step = hp.ChamferStepGraph(B, N, M, dev, with_host_io=True)
    step.xyz1.copy_(a_h.to(dev))
    step.xyz2.copy_(b_h.to(dev))
    step.xyz1_host.copy_(a_h)
    step.xyz2_host.copy_(b_h)

    # forward-only / backward-only graphs for the per-kernel roofline
    fwd_in = (step.xyz1, step.xyz2)
    fwd_graph, fwd_out, _ = hp.graphs._capture(lambda: hp.chamfer_forward(*fwd_in, want_inverse=True), dev)
    one = torch.ones((), device=dev)
    bwd_graph, _bo, _ = hp.graphs._capture(
        lambda: hp.chamfer_backward(step.xyz1, step.xyz2, fwd_out[2], fwd_out[4], one, fwd_out[5]), dev)

    # eager public API (autograd module), for reference: CPU/launch bound at this size
    a = a_h.to(dev).requires_grad_(True)
    b = b_h.to(dev).requires_grad_(True)

    def step_eager():
        a.grad = b.grad = None
        loss = loss_mod(b, a)  # forward(preds, gts)
        loss.backward()

    # --- roofline denominators, measured live on this GPU ---------------------------------
    peak_ffma = native.measure_peak(0, 8192, stream.cuda_stream)
    peak_ffma2 = native.measure_peak(1, 8192, stream.cuda_stream)
    peak_mufu = native.measure_peak(2, 8192, stream.cuda_stream)
    fp32_peak = max(peak_ffma, peak_ffma2)

    # the dominant kernel alone (nn_ring_kernel), on a private workspace that is discarded afterwards (bench library)
    blib = native.load_bench()
    ring_ws = torch.zeros(blib.hp_chamfer_workspace_bytes(B, N, M), dtype=torch.uint8, device=dev)

    def ring_only():
        native.check_bench(blib.hp_measure_chamfer_ring_only(B, N, step.xyz1.data_ptr(), M, step.xyz2.data_ptr(), ring_ws.data_ptr(),
                                                             ring_ws.numel(), stream.cuda_stream), "hp_measure_chamfer_ring_only")

    with ClockSampler(local_rank) as clocks:
        step_ms = _events_timed(torch, step.replay, args.steps, args.warmup, flush, stream, barrier)
        ring_ms = _events_timed(torch, ring_only, args.steps, 3, flush, stream, barrier)
        fwd_ms = _events_timed(torch, fwd_graph.replay, args.steps, 3, flush, stream, barrier)
        bwd_ms = _events_timed(torch, bwd_graph.replay, args.steps, 3, flush, stream, barrier)
    n_e2e = max(20, args.steps // 2)
    # e2e, dependent steps (what a trainer's loop sees): H2D of both clouds + step + D2H of the loss, one after the other
    e2e_ms = _events_timed(torch, step.run_from_host_loss_only, n_e2e, 3, flush, stream, barrier)
    # ... the same as two half-batch steps (the second half's copies under the first half's kernels): an opt-in that does not pay
    split = hp.ChamferStepGraph(B, N, M, dev, with_host_io=True, split_host_io=2)
    split.xyz1_host.copy_(a_h)
    split.xyz2_host.copy_(b_h)
    e2e_split_ms = statistics.mean(_events_timed(torch, split.run_from_host_loss_only, n_e2e, 3, flush, stream, barrier))
    step.run_from_host_loss_only()
    torch.cuda.synchronize(dev)
    gs1, gs2 = split.grad_outputs_on_device()
    gp1, gp2 = step.grad_outputs_on_device()
    assert torch.equal(gs1, gp1) and torch.equal(gs2, gp2), "split and unsplit host graphs differ"
    assert abs(float(split.loss_host) - float(step.loss_host)) <= 2e-6 * abs(float(step.loss_host))
    del split
    # ... and with both gradients copied back as well
    e2e_full_ms = _events_timed(torch, step.run_from_host, n_e2e, 3, flush, stream, barrier)
    # pipelined over INDEPENDENT steps: ChamferHostPipeline (copies of neighbouring steps overlap the compute).  Every step
    # moves fresh data over PCIe, so there is nothing to evict between steps.
    pipe = hp.ChamferHostPipeline(B, N, M, dev, depth=4)
    for _ in range(5):
        pipe.submit(step.xyz1_host, step.xyz2_host)
    pipe.drain()
    n_pipe = max(100, args.steps)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    p0.record(stream)
    pipe.fork_from(stream)          # every slot stream starts after p0
    for _ in range(n_pipe):
        last = pipe.submit(step.xyz1_host, step.xyz2_host)
    pipe.join_into(stream)          # p1 follows the last D2H copy of every slot
    p1.record(stream)
    pipe.result(last)
    pipe.drain()
    barrier()
    pipe_ms = p0.elapsed_time(p1) / n_pipe
    step.run_from_host()
    torch.cuda.synchronize(dev)
    assert torch.equal(pipe.result(last)[1], step.grad_xyz1_host), "pipelined and single-graph e2e paths differ"
    eager_ms = _events_timed(torch, step_eager, max(20, args.steps // 4), 3, flush, stream, barrier)
    # the graph and the eager module must agree bit for bit
    step.replay()
    step_eager()
    torch.cuda.synchronize(dev)
    assert torch.equal(step.grad_xyz1, a.grad) and torch.equal(step.grad_xyz2, b.grad), "graph and eager paths differ"

    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    e2e_total = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_total, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    value = world * PAIRS_PER_STEP * args.steps / total_s
    e2e_value = world * PAIRS_PER_STEP * len(e2e_ms) / (float(e2e_total.item()) * 1e-3)

    peaks_file = {}
    try:
        peaks_file = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    other = comparators = c4 = None
    if rank == 0 and world == 1 and not args.no_other_paths:
        other = _other_paths(torch, hp, dev, fp32_peak, peak_mufu, peaks_file.get("hbm_gbs"), flush, stream, barrier)
        comparators = _secondary_comparators(torch, hp, dev, flush, stream, barrier)
        try:
            c4 = _hypernet_and_c4(torch, hp, dev, flush, stream, barrier)
        except Exception as e:  # the headline line must survive a failure of a side measurement
            c4 = {"error": repr(e)}
    metrics_eval = None
    if not args.no_metrics_eval:
        metrics_eval = _metrics_eval(torch, dist, hp, dev, world, rank, barrier, peak_mufu, fp32_peak, args.metrics_emd_1nn)

    if rank == 0:
        fwd_avg_s = statistics.mean(fwd_ms) * 1e-3
        step_avg_s = total_s / args.steps
        achieved = (PAIRS_PER_STEP * FLOP_PER_PAIR) / fwd_avg_s / 1e12
        ring_avg_s = statistics.mean(ring_ms) * 1e-3
        roofline = {
            "bound": "fp32", "kernel": "nn_ring_kernel: all-pairs squared distances + argmin, both directions (the dominant kernel of the step)",
            "achieved": (PAIRS_PER_STEP * FLOP_PER_PAIR) / ring_avg_s / 1e12, "peak": fp32_peak / 1e12,
            "unit": "TFLOP/s", "frac": (PAIRS_PER_STEP * FLOP_PER_PAIR) / ring_avg_s / fp32_peak,
            "traffic": NCU_RING_DRAM_BYTES,
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of nn_ring_kernel, ncu --set full capture "
                              "profiles/r02_ncu_full_nn_ring.txt (inputs are 1.5 MiB; results stay in L2 for the tail kernel)",
            "peak_source": "hp_measure_peak (libhp_b200_bench.so): register-resident FFMA/FFMA2 chains on all SMs, measured live in this run "
                           "(MEASURED_PEAKS.json has no FP32 entry; K=3 keeps the path off the tensor cores); "
                           "nominal 148*128*2*1.965 GHz = 74.4",
            "algorithmic_flop_per_launch": PAIRS_PER_STEP * FLOP_PER_PAIR,
            "executed_flop_per_launch": PAIRS_PER_STEP * FLOP_PER_PAIR // 2,
            "note": "algorithmic = 8 FLOP per ORDERED (query,candidate) pair (SURVEY 8d); the ring kernel evaluates each "
                    "unordered pair once for both directions, so executed FLOP = half.  The kernel is bound by instruction "
                    "issue (packed fp32x2 ops hold the issue port for two cycles; DESIGN.md 4.1), not by a pipe.",
            "kernel_ms": statistics.mean(ring_ms),
            "forward_ms": statistics.mean(fwd_ms), "forward_frac": achieved / (fp32_peak / 1e12),
            "forward_what": "nn_ring_kernel + nn_ring_unpack_kernel: everything nn_distance returns (distances, indices, loss)",
            "bwd_kernel_ms": statistics.mean(bwd_ms),
            "fwd+bwd_frac": (PAIRS_PER_STEP * FLOP_PER_PAIR) / step_avg_s / fp32_peak,
            "fwd+bwd_what": "the whole step (value): nn_ring_kernel + nn_ring_tail_kernel (per-cloud tickets, runs under the ring kernel's last wave)",
            "peak_ffma_tflops": peak_ffma / 1e12, "peak_ffma2_tflops": peak_ffma2 / 1e12, "peak_mufu_tex2": peak_mufu / 1e12,
            "hbm_peak_gbs_measured": peaks_file.get("hbm_gbs"),
            "metrics_eval": metrics_eval,
        }
        cpu = cpu_c1 = None
        if world == 1 and not args.no_cpu_baseline:
            times, cores, kind = cpu_reference_arm(3, 1, B)
            cpu = {"value": PAIRS_PER_STEP / statistics.mean(times), "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": "3 full steps (B=32, 2048x2048): the reference's pure-torch ChamferLoss fwd+bwd (losses/champfer_loss.py) on CPU",
                   "ms_per_step": statistics.mean(times) * 1e3}
            try:
                cpu_c1 = cpu_reference_c1()
            except Exception as e:
                cpu_c1 = {"error": repr(e)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_s * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": _config(),
            "launch": "step captured once as a CUDA graph (ChamferStepGraph: nn_ring_kernel + nn_ring_tail_kernel, the second launched "
                      "programmatically dependent and gated by per-cloud tickets), one replay per step",
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": step.h2d_bytes, "d2h_bytes_per_step": step.d2h_bytes_loss_only,
                    "ms_per_step": statistics.mean(e2e_ms),
                    "api": "ChamferStepGraph.run_from_host_loss_only: DEPENDENT steps, each = pinned host clouds -> H2D -> fwd+bwd -> D2H of "
                           "the loss; the gradients stay on the device (what the TargetNetwork backward consumes)",
                    "split_in_two_halves_ms_per_step": e2e_split_ms,
                    "split_note": "ChamferStepGraph(split_host_io=2): the second half's copies fly under the first half's kernels; same "
                                  "per-cloud bits.  Opt-in: the extra launches cost about what the overlap gains (graphs._host_io_bounds)",
                    "l2_policy": "L2 flushed between steps like `value`; the inputs arrive over PCIe every step anyway",
                    "with_gradients_d2h_ms_per_step": statistics.mean(e2e_full_ms), "with_gradients_d2h_bytes": step.d2h_bytes,
                    "pipelined_independent_steps_ms_per_step": pipe_ms,
                    "pipelined_api": "ChamferHostPipeline.submit/result (4 buffer sets, one stream each; loss AND both gradients copied "
                                     "back): a THROUGHPUT over independent steps, not the latency of a trainer's dependent loop"},
            "eager_api": {"ms_per_step": statistics.median(eager_ms), "value": PAIRS_PER_STEP / (statistics.median(eager_ms) * 1e-3),
                          "mean_ms_per_step": statistics.mean(eager_ms),
                          "note": "ChamferLoss()(preds, gts); loss.backward() through torch autograd, no graph: CPU launch path bound"},
            "gpu_launches": step.launches_per_replay * args.steps,  # timed `value` region only
            "roofline": roofline,
            "cpu_baseline": cpu,
            "cpu_baseline_c1": cpu_c1,
            "other_paths": other,
            "secondary_comparators": comparators,
            "c4_full_step": c4,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-paths", action="store_true", help="skip the TargetNetwork / EMD / pairwise-CD side measurements")
    ap.add_argument("--no-metrics-eval", action="store_true", help="skip the sharded compute_all_metrics evaluation (config C5)")
    ap.add_argument("--metrics-emd-1nn", action="store_true",
                    help="also run the two 1000x1000 self-distance EMD matrices of 1-NNA-EMD (3x the EMD time of the default run)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
