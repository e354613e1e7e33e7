"""GPU parity tests for the fused TargetNetwork kernels, through the C ABI (ctypes -> libhp_b200.so).
Bars: outputs and weight/point gradients within 1e-5 relative of the oracle (fp32 C forward, float64
numpy backward, both pinned to the reference's model/target_network.py by tests/golden/cpu_reference.npz)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
FAST = [32, 64, 128, 64]


def _rel_err(x, ref):
    x, ref = np.asarray(x, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-30))


def _preacts_f64(w, x, loc, use_bias):
    """float64 pre-activations of every hidden layer: list of [B, N, width]."""
    w, h = np.asarray(w, np.float64), np.asarray(x, np.float64)
    dims = [3] + list(loc) + [3]
    off, pre = 0, []
    for l in range(len(dims) - 2):
        i, o = dims[l], dims[l + 1]
        W = w[:, off:off + i * o].reshape(-1, o, i)
        off += i * o
        z = np.einsum("bni,boi->bno", h, W)
        if use_bias:
            z = z + w[:, off:off + o][:, None, :]
            off += o
        pre.append(z)
        h = np.maximum(z, 0.0)
    return pre


def _inputs(b, n, loc, use_bias, seed, wscale=0.15):
    """Random weights / points / upstream grads.  A point whose pre-activation is within fp32 rounding of zero at some
    hidden unit has an ill-defined ReLU gate (fp32 and the float64 oracle may legitimately disagree, which changes the
    gradient by that point's whole contribution); such points (about one in 10^5) are replaced by a copy of a safe one
    so that the comparison tests arithmetic, not a measure-zero tie."""
    g = torch.Generator().manual_seed(seed)
    from oracle import oracle as O

    W = O.target_network_num_weights(loc, use_bias)
    w = torch.randn(b, W, generator=g) * wscale
    x = torch.randn(b, n, 3, generator=g) * 0.6
    go = torch.randn(b, n, 3, generator=g)
    for _ in range(3):
        pre = _preacts_f64(w.numpy(), x.numpy(), loc, use_bias)
        unsafe = np.zeros((b, n), bool)
        for z in pre:
            unsafe |= (np.abs(z) < 1e-5 * max(1.0, float(np.abs(z).max()))).any(axis=2)
        if not unsafe.any():
            break
        for s in range(b):
            safe = np.flatnonzero(~unsafe[s])
            if len(safe) == 0:
                continue
            x[s, torch.from_numpy(np.flatnonzero(unsafe[s]))] = x[s, int(safe[0])].clone()
    return w, x, go


def test_golden_reference_module(hp, golden_cpu):
    """Vectors produced by the reference's own TargetNetwork class (tests/golden/make_golden_cpu.py)."""
    g = golden_cpu
    w = torch.from_numpy(g["tn_w"]).to(DEV).requires_grad_(True)
    x = torch.from_numpy(g["tn_x"]).to(DEV)
    y = hp.target_network_forward(w, x, FAST, True)
    assert _rel_err(y.detach().cpu().numpy(), g["tn_y"]) < 1e-5
    (y * torch.from_numpy(g["tn_gout"]).to(DEV)).sum().backward()
    assert _rel_err(w.grad.cpu().numpy(), g["tn_grad_w"]) < 1e-5
    # generic path: other widths, no bias
    y2 = hp.target_network_forward(torch.from_numpy(g["tn2_w"]).to(DEV), torch.from_numpy(g["tn2_x"]).to(DEV), [16, 8], False)
    assert _rel_err(y2.cpu().numpy(), g["tn2_y"]) < 1e-5


@pytest.mark.parametrize("b,n", [(1, 1), (2, 127), (3, 128), (2, 129), (5, 300), (4, 2048), (70, 256), (160, 130)])
@pytest.mark.parametrize("use_bias", [True, False])
def test_forward_backward_vs_oracle_fast_shape(hp, oracle, b, n, use_bias):
    w, x, go = _inputs(b, n, FAST, use_bias, seed=b * 100 + n)
    wd = w.to(DEV).requires_grad_(True)
    xd = x.to(DEV).requires_grad_(True)
    y = hp.target_network_forward(wd, xd, FAST, use_bias)
    oy = oracle.target_network_forward(w.numpy(), x.numpy(), FAST, use_bias)
    assert tuple(y.shape) == (b, n, 3)
    assert _rel_err(y.detach().cpu().numpy(), oy) < 1e-5
    (y * go.to(DEV)).sum().backward()
    ogw, ogx = oracle.target_network_backward_f64(w.numpy(), x.numpy(), go.numpy(), FAST, use_bias)
    assert _rel_err(wd.grad.cpu().numpy(), ogw) < 1e-5
    assert _rel_err(xd.grad.cpu().numpy(), ogx) < 1e-5
    # weights-only gradient (the trainer's case: input points carry no grad) must agree with the above
    wd2 = w.to(DEV).requires_grad_(True)
    (hp.target_network_forward(wd2, x.to(DEV), FAST, use_bias) * go.to(DEV)).sum().backward()
    assert torch.equal(wd2.grad, wd.grad)


def test_both_arithmetic_modes_hold_the_bar(hp, oracle):
    """The default (error-compensated 3xTF32: tcgen05 forward, mma.sync backward), the all-mma.sync variant and the fp32 FFMA kernels
    against the same oracle; a sample whose tiles are shared by several CTAs (n = 2500: partial-gradient fold) and a ragged tail tile."""
    b, n = 5, 2500
    w, x, go = _inputs(b, n, FAST, True, seed=77)
    oy = oracle.target_network_forward(w.numpy(), x.numpy(), FAST, True)
    ogw, ogx = oracle.target_network_backward_f64(w.numpy(), x.numpy(), go.numpy(), FAST, True)
    got = {}
    try:
        for mode in ("fp32", "mma.sync", "tf32x3"):
            hp.target_network_set_mode(mode)
            wd = w.to(DEV).requires_grad_(True)
            xd = x.to(DEV).requires_grad_(True)
            y = hp.target_network_forward(wd, xd, FAST, True)
            (y * go.to(DEV)).sum().backward()
            assert _rel_err(y.detach().cpu().numpy(), oy) < 1e-5, mode
            assert _rel_err(wd.grad.cpu().numpy(), ogw) < 1e-5, mode
            assert _rel_err(xd.grad.cpu().numpy(), ogx) < 1e-5, mode
            y2 = hp.target_network_forward(wd.detach(), xd.detach(), FAST, True)
            assert torch.equal(y2, y.detach()), mode  # run-to-run identical (no atomics anywhere)
            got[mode] = y.detach()
    finally:
        hp.target_network_set_mode("tf32x3")
    assert not torch.equal(got["fp32"], got["tf32x3"]) and not torch.equal(got["mma.sync"], got["tf32x3"])  # the switch selects kernels


def test_channels_first_and_shared_cloud(hp, oracle):
    b, n = 6, 515
    w, x, go = _inputs(b, n, FAST, True, seed=9)
    wd = w.to(DEV).requires_grad_(True)
    y_cf = hp.target_network_forward(wd, x.to(DEV), FAST, True, channels_first=True)  # [B,3,N], full_model.py:68,74
    assert tuple(y_cf.shape) == (b, 3, n)
    y = hp.target_network_forward(wd.detach(), x.to(DEV), FAST, True)
    assert torch.equal(y_cf.detach().permute(0, 2, 1), y)
    (y_cf * go.to(DEV).permute(0, 2, 1)).sum().backward()
    ogw, _ = oracle.target_network_backward_f64(w.numpy(), x.numpy(), go.numpy(), FAST, True)
    assert _rel_err(wd.grad.cpu().numpy(), ogw) < 1e-5
    # one shared input cloud for every sample
    xs = x[0].to(DEV).requires_grad_(True)
    wd3 = w.to(DEV).requires_grad_(True)
    ys = hp.target_network_forward(wd3, xs, FAST, True)
    xe = x[:1].expand(b, n, 3).contiguous()
    oy = oracle.target_network_forward(w.numpy(), xe.numpy(), FAST, True)
    assert _rel_err(ys.detach().cpu().numpy(), oy) < 1e-5
    (ys * go.to(DEV)).sum().backward()
    ogw, ogx = oracle.target_network_backward_f64(w.numpy(), xe.numpy(), go.numpy(), FAST, True)
    assert _rel_err(wd3.grad.cpu().numpy(), ogw) < 1e-5
    assert _rel_err(xs.grad.cpu().numpy(), ogx.sum(0)) < 1e-5


@pytest.mark.parametrize("loc,use_bias", [([16, 8], False), ([7], True), ([32, 64, 128, 64, 5], True), ([200, 31], True)])
def test_generic_widths_vs_oracle(hp, oracle, loc, use_bias):
    b, n = 3, 77
    w, x, go = _inputs(b, n, loc, use_bias, seed=len(loc) * 7 + 1, wscale=0.2)
    wd = w.to(DEV).requires_grad_(True)
    xd = x.to(DEV).requires_grad_(True)
    y = hp.target_network_forward(wd, xd, loc, use_bias)
    oy = oracle.target_network_forward(w.numpy(), x.numpy(), loc, use_bias)
    assert _rel_err(y.detach().cpu().numpy(), oy) < 1e-5
    (y * go.to(DEV)).sum().backward()
    ogw, ogx = oracle.target_network_backward_f64(w.numpy(), x.numpy(), go.numpy(), loc, use_bias)
    assert _rel_err(wd.grad.cpu().numpy(), ogw) < 1e-5
    assert _rel_err(xd.grad.cpu().numpy(), ogx) < 1e-5


def test_module_interface_matches_reference_class(hp, oracle):
    """TargetNetwork(config, weights).forward(x): one sample, x [N,3] -> [N,3] (model/target_network.py:31-38)."""
    w, x, _ = _inputs(1, 2048, FAST, True, seed=4)
    cfg = {"use_bias": True, "layer_out_channels": FAST}
    wd = w[0].to(DEV).requires_grad_(True)
    net = hp.TargetNetwork(cfg, wd)
    assert tuple(net.layers["3"]["weight"].shape) == (128, 64) and tuple(net.output["bias"].shape) == (3,)
    y = net(x[0].to(DEV))
    oy = oracle.target_network_forward(w.numpy(), x.numpy(), FAST, True)[0]
    assert tuple(y.shape) == (2048, 3) and _rel_err(y.detach().cpu().numpy(), oy) < 1e-5
    y.sum().backward()
    assert wd.grad is not None and tuple(wd.grad.shape) == (19011,)
    with pytest.raises(AssertionError):
        hp.TargetNetwork(cfg, wd[:-1].detach())


def test_full_size_c4_properties(hp):
    """BASELINE config C4 shape (B=64 x 2048 points): determinism, linearity of the backward in grad_out,
    and batch independence (each sample only sees its own weights)."""
    b, n = 64, 2048
    w, x, go = _inputs(b, n, FAST, True, seed=1)
    wd, xd, gd = w.to(DEV), x.to(DEV), go.to(DEV)

    def fwd_bwd(wt, g_out):
        wt = wt.clone().requires_grad_(True)
        y = hp.target_network_forward(wt, xd, FAST, True)
        y.backward(g_out)
        return y.detach(), wt.grad

    y1, g1 = fwd_bwd(wd, gd)
    y2, g2 = fwd_bwd(wd, gd)
    assert torch.equal(y1, y2) and torch.equal(g1, g2), "fused TargetNetwork must be bitwise reproducible"
    _, g3 = fwd_bwd(wd, 2.0 * gd)
    torch.testing.assert_close(g3, 2.0 * g1, rtol=1e-6, atol=0)  # scaling by 2 is exact in fp32
    perm = torch.randperm(b, generator=torch.Generator().manual_seed(0)).to(DEV)
    yp = hp.target_network_forward(wd[perm], xd[perm], FAST, True)
    assert torch.equal(yp, y1[perm])
    # vs plain torch fp32 (the reference's op sequence) on a few samples
    for s in (0, 17, 63):
        h = xd[s]
        off = 0
        dims = [3] + FAST + [3]
        for l in range(5):
            i, o = dims[l], dims[l + 1]
            Wl = wd[s, off:off + i * o].view(o, i)
            off += i * o
            h = torch.mm(h, Wl.t()) + wd[s, off:off + o]
            off += o
            if l < 4:
                h = torch.relu(h)
        torch.testing.assert_close(y1[s], h, rtol=1e-5, atol=1e-5 * float(h.abs().max()))


def test_errors(hp):
    w = torch.zeros(2, 19011, device=DEV)
    x = torch.zeros(2, 16, 3, device=DEV)
    with pytest.raises(AssertionError):
        hp.target_network_forward(w[:, :-1], x, FAST, True)
    with pytest.raises(RuntimeError):
        hp.target_network_forward(w.cpu(), x, FAST, True)
    with pytest.raises(RuntimeError):
        hp.target_network_forward(w, x.double(), FAST, True)
    with pytest.raises(RuntimeError):
        hp.target_network_forward(w, x[:1], FAST, True)
    y = hp.target_network_forward(w[:0], x[:0], FAST, True)
    assert tuple(y.shape) == (0, 16, 3)


def test_graph_capture_matches_eager(hp):
    """TargetNetworkStepGraph / ChamferStepGraph replay the same kernels: bit-identical to the eager calls."""
    b, n = 8, 640
    tg = hp.TargetNetworkStepGraph(b, n, FAST, True, DEV, channels_first=True)
    w, x, go = _inputs(b, n, FAST, True, seed=21)
    tg.weights.copy_(w.to(DEV))
    tg.points.copy_(x.to(DEV))
    tg.grad_out.copy_(go.to(DEV).permute(0, 2, 1))
    out, gw = tg.replay()
    torch.cuda.synchronize()
    wd = w.to(DEV).requires_grad_(True)
    y = hp.target_network_forward(wd, x.to(DEV), FAST, True, channels_first=True)
    y.backward(go.to(DEV).permute(0, 2, 1).contiguous())
    assert torch.equal(out, y.detach()) and torch.equal(gw, wd.grad)

    cg = hp.ChamferStepGraph(3, 700, 900, DEV, with_host_io=True)
    g = torch.Generator().manual_seed(2)
    a, c = torch.rand(3, 700, 3, generator=g) - 0.5, torch.rand(3, 900, 3, generator=g) - 0.5
    cg.xyz1.copy_(a.to(DEV))
    cg.xyz2.copy_(c.to(DEV))
    loss, g1, g2 = cg.replay()
    torch.cuda.synchronize()
    ad, cd = a.to(DEV).requires_grad_(True), c.to(DEV).requires_grad_(True)
    ref = hp.ChamferLoss()(cd, ad)
    ref.backward()
    assert torch.equal(loss.reshape(()), ref.detach()) and torch.equal(g1, ad.grad) and torch.equal(g2, cd.grad)
    lh, g1h, g2h = cg.run_from_host(a.pin_memory(), c.pin_memory())
    torch.cuda.synchronize()
    assert torch.equal(lh.reshape(()), ref.detach().cpu()) and torch.equal(g1h, ad.grad.cpu()) and torch.equal(g2h, cd.grad.cpu())


@pytest.mark.parametrize("b,split", [(5, True), (4, True), (4, False), (4, None)])
def test_host_graph_loss_only_split_matches_eager(hp, b, split):
    """ChamferStepGraph.run_from_host_loss_only, also as two half-batch steps whose second-half copies fly under the first half's
    kernels (split_host_io): per-cloud results (distances, indices, both gradients) are bit-identical to the eager step -- a cloud's
    numbers do not depend on which launch carried it -- and the loss is the sum of the halves' losses; replays with new inputs
    (zero-restored workspace shared by the two halves) stay exact."""
    n, m = 700, 900
    cg = hp.ChamferStepGraph(b, n, m, DEV, with_host_io=True, split_host_io=split)
    assert cg.host_io_split == bool(split)  # opt-in: None means one part
    g = torch.Generator().manual_seed(31 + b)
    one = torch.ones((), device=DEV)
    for rep in range(3):
        a, c = torch.rand(b, n, 3, generator=g) - 0.5, torch.rand(b, m, 3, generator=g) - 0.5
        if rep == 2:
            c = c * 0.02  # skewed assignment: the tail's general path
        cg.xyz1_host.copy_(a)
        cg.xyz2_host.copy_(c)
        lh = cg.run_from_host_loss_only()
        torch.cuda.synchronize()
        g1, g2 = cg.grad_outputs_on_device()
        loss, _d1, _i1, _d2, _i2, e1, e2 = hp.chamfer_step(a.to(DEV), c.to(DEV), one)
        assert torch.equal(g1, e1) and torch.equal(g2, e2)
        if cg.host_io_split:
            h = b // 2
            la = hp.chamfer_step(a[:h].to(DEV), c[:h].to(DEV), one)[0]
            lb = hp.chamfer_step(a[h:].to(DEV), c[h:].to(DEV), one)[0]
            assert torch.equal(lh, (la + lb).cpu())
            assert torch.equal(cg._s_d1, _d1) and torch.equal(cg._s_i1, _i1) and torch.equal(cg._s_d2, _d2) and torch.equal(cg._s_i2, _i2)
            torch.testing.assert_close(lh, loss.cpu(), rtol=2e-6, atol=0)
        else:
            assert torch.equal(lh, loss.cpu())


def test_chamfer_step_out_buffers(hp):
    b, n, m = 3, 300, 257
    g = torch.Generator().manual_seed(4)
    a, c = (torch.rand(b, n, 3, generator=g) - 0.5).to(DEV), (torch.rand(b, m, 3, generator=g) - 0.5).to(DEV)
    one = torch.ones((), device=DEV)
    ref = hp.chamfer_step(a, c, one)
    big = [torch.zeros(2, 1, device=DEV), torch.zeros(b + 2, n, device=DEV), torch.zeros(b + 2, n, dtype=torch.int32, device=DEV),
           torch.zeros(b + 2, m, device=DEV), torch.zeros(b + 2, m, dtype=torch.int32, device=DEV), torch.zeros(b + 2, n, 3, device=DEV),
           torch.zeros(b + 2, m, 3, device=DEV)]
    out = tuple([big[0][1]] + [t[1:1 + b] for t in big[1:]])
    got = hp.chamfer_step(a, c, one, out=out)
    for r, o, t in zip(ref, got, big):
        assert torch.equal(r, o)
    for t in big[1:]:
        assert not bool(t[0].any()) and not bool(t[-1].any())  # nothing outside the slices was touched
    with pytest.raises(RuntimeError):
        hp.chamfer_step(a, c, one, out=tuple([big[0][1]] + [t[:b + 1] for t in big[1:]]))


def test_host_pipeline_matches_eager(hp):
    """ChamferHostPipeline overlaps H2D / compute / D2H over several buffer sets: every ticket's results must equal the
    eager module on the same inputs, also when slots are reused."""
    b, n, m = 2, 600, 450
    pipe = hp.ChamferHostPipeline(b, n, m, DEV, depth=2)
    g = torch.Generator().manual_seed(12)
    inputs = [((torch.rand(b, n, 3, generator=g) - 0.5).pin_memory(), (torch.rand(b, m, 3, generator=g) - 0.5).pin_memory())
              for _ in range(5)]
    mod = hp.ChamferLoss()
    for i, (a, c) in enumerate(inputs):
        t = pipe.submit(a, c)
        if i >= 1:  # read the previous ticket while the current one is in flight
            loss, g1, g2 = pipe.result(t - 1)
            ad, cd = inputs[i - 1][0].to(DEV).requires_grad_(True), inputs[i - 1][1].to(DEV).requires_grad_(True)
            ref = mod(cd, ad)
            ref.backward()
            assert torch.equal(loss.reshape(()), ref.detach().cpu())
            assert torch.equal(g1, ad.grad.cpu()) and torch.equal(g2, cd.grad.cpu())
    pipe.drain()
    with pytest.raises(RuntimeError):
        pipe.result(0)


def test_reconstruct_batch_vs_reference_full_model_golden(hp, golden_cpu):
    """The batched replacement of the per-sample loop of FullModel.forward (model/full_model.py:67-74) against the reference's own
    FullModel run on CPU (tests/golden/make_golden_cpu.py section 7): same hypernetwork output, the input clouds re-drawn from
    the same global-RNG seed in the same order, reconstruction [B, 3, N] within 1e-5; gradients reach the weights."""
    g = golden_cpu
    B, N, epoch, seed = (int(v) for v in g["fm_meta"])
    w = torch.from_numpy(g["fm_weights"]).to(DEV).requires_grad_(True)
    cfg = {"use_bias": True, "relu_slope": 0.2, "freeze_layers_learning": False, "layer_out_channels": [32, 64, 128, 64]}
    pcfg = {"target_network_input": {"constant": False, "normalization": {"enable": True, "type": "progressive", "epoch": 100}}}
    torch.manual_seed(seed)
    rec = hp.reconstruct_batch(cfg, pcfg, w, N, epoch, DEV)
    assert rec.shape == (B, 3, N) and rec.is_cuda
    assert _rel_err(rec.detach().cpu().numpy(), g["fm_rec"]) < 1e-5
    rec.square().sum().backward()
    assert w.grad is not None and torch.isfinite(w.grad).all() and float(w.grad.abs().sum()) > 0
    # explicit points override the sampling; per-sample TargetNetwork modules give the same clouds
    torch.manual_seed(seed)
    pts = hp.generate_points_batched(pcfg, epoch, B, (N, 3))
    rec2 = hp.reconstruct_batch(cfg, pcfg, w.detach(), N, epoch, DEV, points=pts)
    assert torch.equal(rec2, rec.detach())
    one = hp.TargetNetwork(cfg, w.detach()[1])(pts[1].to(DEV))
    torch.testing.assert_close(one.t(), rec2[1], rtol=1e-6, atol=1e-7)
