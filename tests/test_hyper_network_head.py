"""f4: the fused hypernetwork head (one GEMM whose output is the TargetNetwork weight layout) against the REFERENCE'S OWN
HyperNetwork class (model/hyper_network.py:5-43, loaded unmodified from baseline/_ref or the reference checkout): outputs and
every parameter gradient at 1e-5, state_dict keys unchanged, optimizer steps land in the fused storage.  Pure torch host code,
so the same test runs on CPU here and (marked gpu) on the B200 with cuBLAS."""
import copy
import importlib.util
import os

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.path.join(REPO, "baseline", "_ref"), os.environ.get("HP_REFERENCE_ROOT", "/root/reference")]


def _reference_hypernetwork_class():
    for root in CANDIDATES:
        f = os.path.join(root, "model", "hyper_network.py")
        if os.path.isfile(f):
            spec = importlib.util.spec_from_file_location("_hp_ref_hyper_network", f)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod.HyperNetwork
    pytest.skip("reference model/hyper_network.py not available (stage it with tools/stage_reference.py)")


CFG = {"input_size": 128, "use_bias": True, "relu_slope": 0.2, "target_network_layer_out_channels": [32, 64, 128, 64],
       "target_network_use_bias": True, "target_network_freeze_layers_learning": False}


def _check(hp, device):
    HyperNetwork = _reference_hypernetwork_class()
    torch.manual_seed(1856)
    ref = HyperNetwork(dict(CFG)).to(device)
    ours = copy.deepcopy(ref)
    hp.fuse_hypernetwork_head(ours)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())          # checkpoints stay compatible
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]
    g = torch.Generator().manual_seed(3)
    x = torch.randn(6, 128, generator=g).to(device)
    gout = torch.randn(6, 19011, generator=g).to(device)
    y_ref, y = ref(x), ours(x)
    assert y.shape == (6, 19011) and y.is_contiguous()                              # [B, W]: what hp_target_network_forward reads
    scale = float(y_ref.abs().max())
    assert float((y - y_ref).abs().max()) <= 1e-5 * scale
    (y_ref * gout).sum().backward()
    (y * gout).sum().backward()
    for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
        assert p.grad is not None and p.grad.shape == q.grad.shape, n
        assert float((p.grad - q.grad).abs().max()) <= 1e-5 * max(float(q.grad.abs().max()), 1e-12), n
    # a second backward accumulates like autograd does for any Parameter (the gradient storage is fresh per call)
    (ours(x) * gout).sum().backward()
    for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
        assert float((p.grad - 2 * q.grad).abs().max()) <= 2e-5 * max(float(q.grad.abs().max()), 1e-12), n
    # optimizer steps on the (view) Parameters are what the next fused forward uses
    opt_o, opt_r = torch.optim.Adam(ours.parameters(), lr=1e-3), torch.optim.Adam(ref.parameters(), lr=1e-3)
    for p in ref.parameters():
        p.grad = p.grad * 2
    opt_o.step()
    opt_r.step()
    y_ref2, y2 = ref(x), ours(x)
    assert float((y_ref2 - y_ref).abs().max()) > 1e-4 * scale                        # the step did something
    assert float((y2 - y_ref2).abs().max()) <= 2e-5 * float(y_ref2.abs().max())
    # moving the Parameters off the fused storage is detected, not silently ignored
    ours.output[0].weight.data = ours.output[0].weight.data.clone()
    with pytest.raises(RuntimeError):
        ours(x)


def test_fused_head_matches_reference_hypernetwork_cpu(hp):
    _check(hp, "cpu")


@pytest.mark.gpu
def test_fused_head_matches_reference_hypernetwork_gpu(hp):
    _check(hp, "cuda:0")
