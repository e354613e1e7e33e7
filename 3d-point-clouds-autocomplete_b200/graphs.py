"""CUDA-graph capture of the hot-path steps.

At the reference's training shapes one Chamfer forward+backward is ~70 us of GPU work, less than the
Python/launch path that enqueues it, so the eager API is CPU-bound.  These helpers capture a step for
fixed shapes ONCE into a CUDA graph (torch.cuda.CUDAGraph; our kernels are launched on the capturing
stream through the C ABI like any other stream) and replay it with one launch:

    step = ChamferStepGraph(batch, n, m, device)          # capture
    step.xyz1.copy_(a); step.xyz2.copy_(b)                # refresh the static inputs (device tensors)
    step.replay()                                         # loss, grad_xyz1, grad_xyz2 are refreshed in place
    step.run_from_host(a_pinned, b_pinned)                # H2D + step + D2H of (loss, grads), one graph

Semantics are those of ``loss = ChamferLoss()(preds=xyz2, gts=xyz1); loss.backward()`` (losses/champfer_loss.py:11-17
as called by core/epoch_loops.py:25-26) with the kernels of chamfer.py; results are bit-identical to the eager path.
"""
from __future__ import annotations

import torch

from ._glue import workspace_owner
from .chamfer import chamfer_step, chamfer_step_supported
from .target_network import target_network_backward, target_network_forward, target_network_num_weights


class _OwnedGraph:
    """A captured graph together with the zero-restored workspaces its kernels point at (kept alive with it)."""

    def __init__(self, graph, workspaces):
        self.graph, self.workspaces = graph, workspaces

    def replay(self):
        self.graph.replay()


def _capture(fn, device, warmup: int = 3):
    """Warm up on a side stream (allocates workspaces, sets kernel attributes), then capture `fn` once.  Workspaces
    requested during warm-up and capture belong to the returned graph object (``_glue.workspace_owner``), not to the
    shared per-stream cache."""
    store = {}
    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream(device))
    with workspace_owner(store):
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            out = fn()
    return _OwnedGraph(graph, store), out, side


class ChamferStepGraph:
    """Chamfer forward (both directions + loss) and backward (both gradients, upstream grad 1) as one graph."""

    def __init__(self, batch: int, n: int, m: int, device, with_host_io: bool = False, split_host_io=None):
        """``split_host_io`` (only with ``with_host_io``): run ``run_from_host_loss_only`` as consecutive part-batch steps so that
        the next part's host->device copies fly under the current part's kernels (default: one part; see ``_host_io_bounds``
        for the measurement that made it an opt-in)."""
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            # both inputs live in ONE allocation (second cloud set 16-byte aligned behind the first): with host I/O the pinned
            # staging buffer has the same layout and a step's inputs arrive with a single host->device copy
            off2 = (batch * n * 3 + 3) & ~3
            self._xyz_both = torch.zeros(off2 + batch * m * 3, device=self.device)
            self.xyz1 = self._xyz_both[:batch * n * 3].view(batch, n, 3)
            self.xyz2 = self._xyz_both[off2:].view(batch, m, 3)
            # any finite placeholder works for capture; callers overwrite xyz1 / xyz2 before replaying
            self.xyz1.uniform_(-0.5, 0.5)
            self.xyz2.uniform_(-0.5, 0.5)
            self._one = torch.ones((), device=self.device)

            def step():
                return chamfer_step(self.xyz1, self.xyz2, self._one)

            self.graph, outs, self._stream = _capture(step, self.device)
            self.loss, self.dist1, self.idx1, self.dist2, self.idx2, self.grad_xyz1, self.grad_xyz2 = outs
            # ring forward + fused tail (unpack, loss, inverse maps, both gradients); big clouds: ring + unpack + gather
            self.launches_per_replay = 2 if chamfer_step_supported(batch, n, m) else 3

            self.host_graph = None
            if with_host_io:
                self._xyz_both_host = torch.zeros(self._xyz_both.numel()).pin_memory()
                self.xyz1_host = self._xyz_both_host[:batch * n * 3].view(batch, n, 3)
                self.xyz2_host = self._xyz_both_host[off2:].view(batch, m, 3)
                self.loss_host = torch.empty(1).pin_memory()
                self.grad_xyz1_host = torch.empty(batch, n, 3).pin_memory()
                self.grad_xyz2_host = torch.empty(batch, m, 3).pin_memory()
                self.xyz1_host.copy_(self.xyz1)
                self.xyz2_host.copy_(self.xyz2)

                def host_step():
                    self._xyz_both.copy_(self._xyz_both_host, non_blocking=True)
                    loss, _d1, _i1, _d2, _i2, g1, g2 = chamfer_step(self.xyz1, self.xyz2, self._one)
                    self.loss_host.copy_(loss, non_blocking=True)
                    self.grad_xyz1_host.copy_(g1, non_blocking=True)
                    self.grad_xyz2_host.copy_(g2, non_blocking=True)
                    return loss, g1, g2

                self.host_graph, self._host_outs, _ = _capture(host_step, self.device)
                self.h2d_bytes = (self.xyz1.numel() + self.xyz2.numel()) * 4
                self.d2h_bytes = (self.xyz1.numel() + self.xyz2.numel()) * 4 + 4

                def host_step_loss_only():
                    # what a trainer's step moves: inputs in, the loss value out; the gradients stay on the device for the
                    # TargetNetwork backward (core/epoch_loops.py:19-39 reads only .item() of the losses)
                    self._xyz_both.copy_(self._xyz_both_host, non_blocking=True)
                    loss, _d1, _i1, _d2, _i2, g1, g2 = chamfer_step(self.xyz1, self.xyz2, self._one)
                    self.loss_host.copy_(loss, non_blocking=True)
                    return loss, g1, g2

                self.host_io_parts = self._host_io_bounds(batch, n, m, split_host_io)  # [0, b1, ..., batch]
                self.host_io_split = len(self.host_io_parts) > 2
                if self.host_io_split:
                    # The step is PCIe-bound from the host (C2: 1.5 MB in = 45-60 us before a 58 us step).  Clouds are independent
                    # and the loss is a sum over clouds, so the batch runs as consecutive part-batch steps: the copies of part k+1
                    # are a parallel branch of the graph (own stream during capture) that flies under the kernels of part k.
                    # Results land in batch slices of full-size buffers; the loss is the sum of the parts' losses (each part in
                    # the library's fixed order, then one fp32 reduction over the parts).
                    dev, bounds = self.device, self.host_io_parts
                    nparts = len(bounds) - 1
                    self._s_loss = torch.empty(nparts, 1, device=dev)
                    self._s_loss_sum = torch.empty(1, device=dev)
                    self._s_d1, self._s_i1 = torch.empty(batch, n, device=dev), torch.empty(batch, n, dtype=torch.int32, device=dev)
                    self._s_d2, self._s_i2 = torch.empty(batch, m, device=dev), torch.empty(batch, m, dtype=torch.int32, device=dev)
                    self._s_g1, self._s_g2 = torch.empty(batch, n, 3, device=dev), torch.empty(batch, m, 3, device=dev)
                    self._copy_stream = torch.cuda.Stream(device=dev)

                    def copy_in(lo, hi):
                        self.xyz1[lo:hi].copy_(self.xyz1_host[lo:hi], non_blocking=True)
                        self.xyz2[lo:hi].copy_(self.xyz2_host[lo:hi], non_blocking=True)

                    def part(k):
                        lo, hi = bounds[k], bounds[k + 1]
                        outs = (self._s_loss[k], self._s_d1[lo:hi], self._s_i1[lo:hi], self._s_d2[lo:hi], self._s_i2[lo:hi],
                                self._s_g1[lo:hi], self._s_g2[lo:hi])
                        chamfer_step(self.xyz1[lo:hi], self.xyz2[lo:hi], self._one, out=outs)

                    def host_step_loss_only_split():
                        main = torch.cuda.current_stream(dev)
                        copy_in(bounds[0], bounds[1])
                        arrived = [torch.cuda.Event() for _ in range(nparts)]
                        arrived[0].record(main)
                        with torch.cuda.stream(self._copy_stream):  # later parts: behind the first part's copies, beside its kernels
                            self._copy_stream.wait_event(arrived[0])
                            for k in range(1, nparts):
                                copy_in(bounds[k], bounds[k + 1])
                                arrived[k].record(self._copy_stream)
                        part(0)
                        for k in range(1, nparts):
                            main.wait_event(arrived[k])
                            part(k)
                        main.wait_stream(self._copy_stream)
                        torch.sum(self._s_loss, dim=0, out=self._s_loss_sum)
                        self.loss_host.copy_(self._s_loss_sum, non_blocking=True)
                        return self._s_loss_sum, self._s_g1, self._s_g2

                    self.host_graph_loss, self._host_loss_outs, _ = _capture(host_step_loss_only_split, self.device)
                else:
                    self.host_graph_loss, self._host_loss_outs, _ = _capture(host_step_loss_only, self.device)
                self.d2h_bytes_loss_only = 4

    def _host_io_bounds(self, batch: int, n: int, m: int, split):
        """Batch boundaries of the parts ``run_from_host_loss_only`` runs one after the other.  ``split``: None / False / 0 / 1 = one
        part (the default); True = two halves; an int P = P near-equal parts; a sequence = explicit boundaries.  Every part must be
        a shape of the fused step.  Measured at C2 on a B200 (``tools/time_host_split.py``, ``profiles/r02_host_split.txt``): the two
        copies take 41 us, the step 58 us, one part 100 us at best; two halves 109 us, three parts 120 us -- every extra part costs
        8-10 us (two more launches on grids of less than a wave, the cross-branch dependencies of the graph) and the copies do not
        hide fully, so the split is an opt-in, not the default."""
        if split is None or split is False:
            split = 1
        elif split is True:
            split = 2
        if isinstance(split, (list, tuple)):
            bounds = [int(v) for v in split]
            if bounds[0] != 0 or bounds[-1] != batch or any(b1 <= b0 for b0, b1 in zip(bounds, bounds[1:])):
                raise ValueError(f"split_host_io boundaries must ascend from 0 to {batch}: {bounds}")
        else:
            parts = max(1, min(int(split), batch))
            bounds = [(batch * k) // parts for k in range(parts + 1)]
        if len(bounds) > 2 and not all(chamfer_step_supported(b1 - b0, n, m) for b0, b1 in zip(bounds, bounds[1:])):
            bounds = [0, batch]
        return bounds

    def replay(self):
        """Inputs: self.xyz1 / self.xyz2 (device).  Outputs refreshed in place: loss [1], dist*, idx*, grad_xyz*."""
        self.graph.replay()
        return self.loss, self.grad_xyz1, self.grad_xyz2

    def run_from_host(self, xyz1_host: torch.Tensor = None, xyz2_host: torch.Tensor = None):
        """Pinned-host inputs -> (loss_host, grad_xyz1_host, grad_xyz2_host) pinned-host outputs.  The copies are
        nodes of the graph; call torch.cuda.synchronize() (or record an event) before reading the outputs."""
        if self.host_graph is None:
            raise RuntimeError("construct ChamferStepGraph(with_host_io=True) to use run_from_host")
        if xyz1_host is not None and xyz1_host is not self.xyz1_host:
            self.xyz1_host.copy_(xyz1_host)
        if xyz2_host is not None and xyz2_host is not self.xyz2_host:
            self.xyz2_host.copy_(xyz2_host)
        self.host_graph.replay()
        return self.loss_host, self.grad_xyz1_host, self.grad_xyz2_host

    def run_from_host_loss_only(self):
        """Pinned-host inputs (``xyz1_host`` / ``xyz2_host``) -> ``loss_host``; the gradients stay on the device
        (``grad_outputs_on_device()``).  One graph: ONE H2D copy of both cloud sets, ring kernel, tail kernel, one 4-byte D2H
        copy -- or part-batch steps with the next part's copies under the current part's kernels when ``host_io_split``."""
        if self.host_graph is None:
            raise RuntimeError("construct ChamferStepGraph(with_host_io=True) to use run_from_host_loss_only")
        self.host_graph_loss.replay()
        return self.loss_host

    def grad_outputs_on_device(self):
        return self._host_loss_outs[1], self._host_loss_outs[2]


class TargetNetworkStepGraph:
    """Fused TargetNetwork forward + backward w.r.t. the weights for fixed (batch, n) as one graph
    (the model/full_model.py:67-74 loop and its autograd, per training step).  Static inputs: ``weights``,
    ``points``, ``grad_out``; outputs refreshed in place: ``out``, ``grad_weights``."""

    def __init__(self, batch: int, n: int, layer_out_channels, use_bias: bool, device, channels_first: bool = True):
        self.device = torch.device(device)
        W = target_network_num_weights(layer_out_channels, use_bias)
        loc = tuple(int(c) for c in layer_out_channels)
        with torch.cuda.device(self.device):
            self.weights = torch.randn(batch, W, device=self.device) * 0.1
            self.points = torch.rand(batch, n, 3, device=self.device) - 0.5
            self.grad_out = torch.randn((batch, 3, n) if channels_first else (batch, n, 3), device=self.device)

            def step():
                out = target_network_forward(self.weights, self.points, loc, use_bias, channels_first)
                gw, _ = target_network_backward(self.weights, self.points, self.grad_out, loc, use_bias, channels_first)
                return out, gw

            self.graph, (self.out, self.grad_weights), self._stream = _capture(step, self.device)

    def replay(self):
        self.graph.replay()
        return self.out, self.grad_weights


class HotPathStepGraph:
    """The point-set hot path of one HyperPocket training step (BASELINE config C4 without the encoder / hypernetwork):

        rec  = TargetNetwork_b(points_b)            for every sample b     (model/full_model.py:67-74)
        loss = loss_coef * ChamferLoss(gt, rec)                            (core/epoch_loops.py:25-26)
        d loss / d weights                                                  (what autograd hands the hypernetwork)

    as one CUDA graph: fused TargetNetwork forward -> ring Chamfer forward -> fused Chamfer tail (unpack + loss + both
    gradients) -> fused TargetNetwork backward.  Static inputs: ``weights`` [B,W], ``points`` [B,N,3], ``gt`` [B,N,3]; outputs refreshed in
    place: ``rec`` [B,N,3], ``loss`` [1] (unscaled Chamfer sum), ``grad_weights`` [B,W]."""

    def __init__(self, batch: int, n: int, layer_out_channels, use_bias: bool, device, loss_coef: float = 0.05):
        self.device = torch.device(device)
        W = target_network_num_weights(layer_out_channels, use_bias)
        loc = tuple(int(c) for c in layer_out_channels)
        with torch.cuda.device(self.device):
            self.weights = torch.randn(batch, W, device=self.device) * 0.1
            self.points = torch.rand(batch, n, 3, device=self.device) - 0.5
            self.gt = torch.rand(batch, n, 3, device=self.device) - 0.5
            self._coef = torch.full((), float(loss_coef), device=self.device)

            def step():
                rec = target_network_forward(self.weights, self.points, loc, use_bias, False)
                loss, _d1, _i1, _d2, _i2, _g_gt, g_rec = chamfer_step(self.gt, rec, self._coef)
                gw, _ = target_network_backward(self.weights, self.points, g_rec, loc, use_bias, False)
                return rec, loss, gw

            self.graph, (self.rec, self.loss, self.grad_weights), self._stream = _capture(step, self.device)
            self.launches_per_replay = 4 if chamfer_step_supported(batch, n, n) else 5

    def replay(self):
        self.graph.replay()
        return self.rec, self.loss, self.grad_weights


class FullModelStepGraph:
    """ONE whole training step of the reference's trainer (core/epoch_loops.py:14-39; BASELINE config C4) as ONE CUDA graph:

        latent            = full_model.mode.get_latent(...)          stock encoder(s)              (model/full_model.py:65)
        weights [B,19011] = full_model.hyper_network(latent)         stock trunk + (fused) head    (:67)
        rec     [B,N,3]   = fused TargetNetwork(weights, points)     one kernel                    (:70-74)
        loss_r            = mean(loss_coef * ChamferLoss(gt, rec))   ring kernel + sectioned tail  (epoch_loops.py:25-26)
        loss_all          = loss_r (+ KLD / B for the generative mode, :28-31)
        loss_all.backward()                                          fused TargetNetwork backward, autograd through the
                                                                     hypernetwork and the encoders
        optimizer.step()                                             (optional; needs a capturable optimizer)

    ``full_model`` is the caller's FullModel (the reference's class or the drop-in) already on ``device`` and in train mode; its
    Parameters are updated in place by every replay when an optimizer is given.  Static inputs, refreshed by the caller before
    ``replay()``: ``existing`` [B,3,Ne] and ``missing`` [B,3,Nm] (the layout FullModel.forward leaves them in, :56-60), ``gt``
    [B,N,3], ``points`` [B,N,3] (the TargetNetwork inputs: ``load_points`` draws them on the host in the reference's RNG order
    and copies them in).  Outputs refreshed in place: ``rec`` [B,N,3] (``rec.permute(0,2,1)`` is the trainer's
    [B,3,N] view), ``loss_r``, ``loss_all``.  Inside the graph nothing is transposed or re-laid-out: the TargetNetwork writes the
    [B,N,3] layout the Chamfer kernels read."""

    def __init__(self, full_model, optimizer, batch: int, n_existing: int, n_missing: int, n_gt: int, device, loss_coef: float = 0.05,
                 warmup: int = 3):
        from .chamfer import ChamferLoss

        self.device = torch.device(device)
        self.model, self.optimizer, self.loss_coef = full_model, optimizer, float(loss_coef)
        tcfg = full_model.target_network_config
        self._loc, self._use_bias = tuple(int(c) for c in tcfg["layer_out_channels"]), bool(tcfg["use_bias"])
        self._pg_cfg = full_model.point_generator_config
        self._loss_fn = ChamferLoss()
        self.generative = bool(full_model.mode.has_generativity())
        with torch.cuda.device(self.device):
            self.existing = torch.rand(batch, 3, n_existing, device=self.device) - 0.5
            self.missing = torch.rand(batch, 3, max(n_missing, 1), device=self.device) - 0.5
            self.gt = torch.rand(batch, n_gt, 3, device=self.device) - 0.5
            self.points = torch.rand(batch, n_gt, 3, device=self.device) - 0.5
            self.points_host = torch.empty(batch, n_gt, 3).pin_memory()
            store = {}
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with workspace_owner(store):
                with torch.cuda.stream(side):
                    for _ in range(warmup):  # also lets Adam create its (capturable) state and cuBLAS pick its kernels
                        self._step()
                torch.cuda.current_stream(self.device).wait_stream(side)
                torch.cuda.synchronize(self.device)
                if optimizer is not None:
                    optimizer.zero_grad(set_to_none=True)
                else:
                    for p in full_model.parameters():
                        p.grad = None
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    self.rec, self.loss_r, self.loss_all = self._step()
            self.graph, self._stream = _OwnedGraph(graph, store), side

    def _step(self):
        m = self.model
        if self.optimizer is not None:
            self.optimizer.zero_grad(set_to_none=True)
        latent, mu, logvar = m.mode.get_latent(m, self.existing, self.missing, None)
        weights = m.hyper_network(latent)
        rec = target_network_forward(weights, self.points, self._loc, self._use_bias, False)
        loss_r = torch.mean(self.loss_coef * self._loss_fn(self.gt, rec))   # ChamferLoss(preds, gts) is symmetric in its arguments
        loss_all = loss_r
        if self.generative:
            kld = 0.5 * (torch.exp(logvar) + torch.square(mu) - 1 - logvar).sum()
            loss_all = loss_r + torch.div(kld, self.existing.shape[0])
        loss_all.backward()
        if self.optimizer is not None:
            self.optimizer.step()
        return rec.detach(), loss_r.detach(), loss_all.detach()

    def load_points(self, epoch: int):
        """Draw this step's TargetNetwork input clouds on the host (global torch CPU RNG, the reference's order) and copy them
        into the static ``points`` tensor (one pinned H2D copy, stream-ordered before the next replay)."""
        from .target_network import generate_points

        for j in range(self.points_host.size(0)):
            self.points_host[j] = generate_points(self._pg_cfg, epoch, (self.points_host.size(1), 3))
        self.points.copy_(self.points_host, non_blocking=True)

    def replay(self):
        self.graph.replay()
        return self.rec, self.loss_r, self.loss_all


class ChamferHostPipeline:
    """Streams Chamfer steps from pinned host memory.  Every slot (buffer set + captured step graph) has its OWN stream that
    carries its H2D copies, its step and its D2H copies in order; ``depth`` slots are in flight, so while slot i computes,
    slot i+1's inputs are copied in and slot i-1's results are copied out by the copy engines, and the small latency-bound
    tail kernel of one step shares the GPU with the ring kernel of the next.  PCIe moves ~1.5 MB each way per step at the
    reference's training shape (about as long as the step itself), and one submit costs the host ~8 driver calls, so the
    end-to-end rate is that of the kernels (measured at B=32, 2048^2: 53.7 us per step with 4 slots, 59.7 with 3, 82.6 with
    2; ``ChamferStepGraph.run_from_host`` serialises the same work: 140 us).

        pipe = ChamferHostPipeline(batch, n, m, device)
        t = pipe.submit(a_pinned, b_pinned)          # returns a ticket immediately
        loss, ga, gb = pipe.result(t)                # pinned host tensors, valid until the slot is reused
                                                     # (``depth`` submissions later)
    """

    def __init__(self, batch: int, n: int, m: int, device, depth: int = 4):
        self.device = torch.device(device)
        self.depth = depth
        with torch.cuda.device(self.device):
            self.slots = []
            for _ in range(depth):
                g = ChamferStepGraph(batch, n, m, self.device)
                slot = {
                    "graph": g,
                    "loss_host": torch.empty(1).pin_memory(),
                    "g1_host": torch.empty(batch, n, 3).pin_memory(),
                    "g2_host": torch.empty(batch, m, 3).pin_memory(),
                    "stream": torch.cuda.Stream(device=self.device),
                    "done": torch.cuda.Event(),
                }
                slot["done"].record(slot["stream"])
                self.slots.append(slot)
        self._count = 0
        self.h2d_bytes = (batch * n * 3 + batch * m * 3) * 4
        self.d2h_bytes = self.h2d_bytes + 4

    def submit(self, xyz1_host: torch.Tensor, xyz2_host: torch.Tensor) -> int:
        """Enqueue one step on pinned host inputs [B,N,3] / [B,M,3]; returns a ticket for ``result``.  The inputs must stay
        untouched until the ticket's result is available."""
        t = self._count
        slot = self.slots[t % self.depth]
        g = slot["graph"]
        with torch.cuda.stream(slot["stream"]):   # stream order: previous use of this slot -> H2D -> step -> D2H
            g.xyz1.copy_(xyz1_host, non_blocking=True)
            g.xyz2.copy_(xyz2_host, non_blocking=True)
            g.graph.replay()
            slot["loss_host"].copy_(g.loss, non_blocking=True)
            slot["g1_host"].copy_(g.grad_xyz1, non_blocking=True)
            slot["g2_host"].copy_(g.grad_xyz2, non_blocking=True)
            slot["done"].record()
        self._count += 1
        return t

    def result(self, ticket: int):
        """(loss [1], grad_xyz1 [B,N,3], grad_xyz2 [B,M,3]) of submission ``ticket`` as pinned host tensors."""
        if ticket < self._count - self.depth or ticket >= self._count:
            raise RuntimeError(f"ticket {ticket} is no longer (or not yet) held by the pipeline")
        slot = self.slots[ticket % self.depth]
        slot["done"].synchronize()
        return slot["loss_host"], slot["g1_host"], slot["g2_host"]

    def fork_from(self, stream=None):
        """Make every slot stream wait for the work enqueued so far on ``stream`` (default: the current stream) --
        e.g. a CUDA event recorded there that opens a timed region."""
        ev = torch.cuda.Event()
        ev.record(stream if stream is not None else torch.cuda.current_stream(self.device))
        for slot in self.slots:
            slot["stream"].wait_event(ev)

    def join_into(self, stream=None):
        """Make ``stream`` (default: the current stream) wait for everything submitted so far."""
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        for slot in self.slots:
            st.wait_event(slot["done"])

    def drain(self):
        for slot in self.slots:
            slot["stream"].synchronize()
