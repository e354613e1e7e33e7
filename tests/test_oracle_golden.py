"""CPU: pin the oracle (oracle/) against golden vectors generated from the reference's own
Python modules (tests/golden/make_golden_cpu.py) and against first principles."""
import numpy as np
import pytest


def test_nn_lattice_matches_reference_torch_min(oracle, golden_cpu):
    g = golden_cpu
    d1, i1, d2, i2 = oracle.nn_distance(g["lat_a"], g["lat_b"])
    # lattice coordinates k/8: expansion form and direct form are both exact -> bit parity
    assert np.array_equal(i1, g["lat_idx_a"]), "argmin (lowest index on ties) differs from torch.min"
    assert np.array_equal(i2, g["lat_idx_b"])
    assert np.array_equal(d1, g["lat_dist_a"])
    assert np.array_equal(d2, g["lat_dist_b"])
    # the lattice has real ties: make sure the test exercises them
    P = g["lat_P"]
    ties = (P == P.min(axis=2, keepdims=True)).sum(axis=2)
    assert (ties > 1).sum() > 10
    assert np.float32(d1.sum(dtype=np.float64) + d2.sum(dtype=np.float64)) == pytest.approx(float(g["lat_loss"]), rel=1e-6)


def test_nn_uniform_matches_reference(oracle, golden_cpu):
    g = golden_cpu
    d1, i1, d2, i2 = oracle.nn_distance(g["uni_a"], g["uni_b"])
    # expansion-form rounding (~1e-7 abs) can only flip an index where the top-2 gap is tiny
    assert (i1 == g["uni_idx_a"]).mean() > 0.99 and (i2 == g["uni_idx_b"]).mean() > 0.99
    np.testing.assert_allclose(d1, g["uni_dist_a"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(d2, g["uni_dist_b"], rtol=0, atol=2e-6)
    loss = d1.sum(dtype=np.float64) + d2.sum(dtype=np.float64)
    assert loss == pytest.approx(float(g["uni_loss"]), rel=1e-5)


@pytest.mark.parametrize("pre", ["lat", "uni"])
def test_nn_grad_matches_reference_autograd(oracle, golden_cpu, pre):
    g = golden_cpu
    a, b = g[f"{pre}_a"], g[f"{pre}_b"]
    d1, i1, d2, i2 = oracle.nn_distance(a, b)
    if pre == "uni":  # use the reference's own argmins so a flipped near-tie cannot matter
        i1, i2 = g["uni_idx_a"], g["uni_idx_b"]
    ga, gb = oracle.nn_distance_grad(a, b, i1, i2, np.ones_like(d1), np.ones_like(d2))
    np.testing.assert_allclose(ga, g[f"{pre}_grad_a"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(gb, g[f"{pre}_grad_b"], rtol=1e-4, atol=2e-6)


def test_nn_first_principles_f64(oracle):
    rng = np.random.default_rng(3)
    a = (rng.random((2, 200, 3), dtype=np.float32) - 0.5)
    b = (rng.random((2, 777, 3), dtype=np.float32) - 0.5)  # > 512 candidates: crosses the reference's tile edge
    d1, i1, d2, i2 = oracle.nn_distance(a, b)
    D = ((a.astype(np.float64)[:, :, None, :] - b.astype(np.float64)[:, None, :, :]) ** 2).sum(-1)
    assert np.array_equal(i1, D.argmin(2)) and np.array_equal(i2, D.argmin(1))
    np.testing.assert_allclose(d1, D.min(2), rtol=1e-6)


def test_nn_ties_lowest_index(oracle):
    a = np.zeros((1, 3, 3), np.float32)
    b = np.zeros((1, 1030, 3), np.float32)
    b[0, :, 0] = 1.0
    b[0, [5, 600, 1029], 0] = 0.5  # three equal minima, two of them in later 512-tiles
    d1, i1, d2, i2 = oracle.nn_distance(a, b)
    assert (i1 == 5).all() and (d1 == 0.25).all()
    assert (i2 == 0).all()  # all a identical -> lowest index


def test_chamfer_torch_port_equals_reference(oracle, golden_cpu):
    import torch

    g = golden_cpu
    loss = oracle.chamfer_loss_torch(torch.from_numpy(g["uni_b"]), torch.from_numpy(g["uni_a"]))
    assert float(loss) == float(g["uni_loss"])
    P = oracle.batch_pairwise_dist_torch(torch.from_numpy(g["lat_a"]), torch.from_numpy(g["lat_b"]))
    assert np.array_equal(P.numpy(), g["lat_P"])


def test_target_network_matches_reference(oracle, golden_cpu):
    g = golden_cpu
    y = oracle.target_network_forward(g["tn_w"], g["tn_x"], [32, 64, 128, 64], True)
    np.testing.assert_allclose(y, g["tn_y"], rtol=1e-5, atol=1e-6)
    y64, _ = oracle.target_network_forward_f64(g["tn_w"], g["tn_x"], [32, 64, 128, 64], True)
    np.testing.assert_allclose(y64, g["tn_y"], rtol=1e-4, atol=1e-5)
    gw, _gx = oracle.target_network_backward_f64(g["tn_w"], g["tn_x"], g["tn_gout"], [32, 64, 128, 64], True)
    np.testing.assert_allclose(gw, g["tn_grad_w"], rtol=1e-3, atol=1e-4)
    y2 = oracle.target_network_forward(g["tn2_w"], g["tn2_x"], [16, 8], False)
    np.testing.assert_allclose(y2, g["tn2_y"], rtol=1e-5, atol=1e-6)
    assert oracle.target_network_num_weights([32, 64, 128, 64], True) == 19011


def test_mmd_cov_and_knn_match_reference(oracle, golden_cpu):
    g = golden_cpu
    r = oracle.mmd_cov(g["mc_M"])
    assert float(r["mmd(Fidelity)"]) == pytest.approx(float(g["mc_mmd"]), rel=1e-6)
    assert float(r["cov(Coverage)"]) == pytest.approx(float(g["mc_cov"]), rel=1e-6)
    assert float(r["mmd_smp"]) == pytest.approx(float(g["mc_mmd_smp"]), rel=1e-6)
    k = oracle.knn(g["knn_Mxx"], g["knn_Mxy"], g["knn_Myy"], 1)
    for key in ("tp", "fp", "fn", "tn", "precision", "recall", "acc_t", "acc_f", "acc"):
        assert k[key] == pytest.approx(float(g["knn_" + key]), rel=1e-6), key


def test_emd_oracle_properties(oracle):
    """No reference vectors exist for EMD on CPU (CUDA-only); check invariants of the auction:
    every row of `match` distributes at most its mass, total matched mass ~ n, cost > 0, and
    identical clouds give (near) zero cost."""
    rng = np.random.default_rng(5)
    a = (rng.random((2, 128, 3), dtype=np.float32) - 0.5)
    b = (rng.random((2, 128, 3), dtype=np.float32) - 0.5)
    match, temp = oracle.approx_match(a, b)
    assert match.shape == (2, 128, 128) and temp.shape == (2, 512)
    assert (match >= 0).all()
    np.testing.assert_allclose(match.sum(axis=(1, 2)), 128.0, rtol=2e-3)
    assert (match.sum(axis=1) <= 1.0 + 1e-4).all() and (match.sum(axis=2) <= 1.0 + 1e-4).all()
    cost = oracle.match_cost_from_match(a, b, match)
    brute = (match * np.sqrt(((a[:, None, :, :] - b[:, :, None, :]) ** 2).sum(-1))).sum(axis=(1, 2))
    np.testing.assert_allclose(cost, brute, rtol=1e-5)
    same = oracle.match_cost(a, a)
    assert (same < 0.02 * cost).all()
    g1, g2 = oracle.match_cost_grad(a, b, match)
    # finite differences of cost with match frozen
    eps = 1e-3
    a2 = a.copy()
    a2[0, 7, 1] += eps
    c2 = oracle.match_cost_from_match(a2, b, match)
    assert (c2[0] - cost[0]) / eps == pytest.approx(g1[0, 7, 1], rel=5e-2, abs=5e-3)


def test_generate_points_batched_matches_reference_sampler(golden_cpu):
    """Host-side input sampling (utils/points.py:8-36): same global-RNG draw order, bit-identical clouds."""
    import importlib

    import torch

    tn = importlib.import_module("3d-point-clouds-autocomplete_b200.target_network")
    cfg = {"target_network_input": {"normalization": {"enable": True, "type": "progressive", "epoch": 100}}}
    for ep in (1, 37, 100, 250):
        torch.manual_seed(1856)
        got = tn.generate_points_batched(cfg, ep, 3, (96, 3), pin=False)
        assert np.array_equal(got.numpy(), golden_cpu[f"gp_ep{ep}"]), ep
    cfg["target_network_input"]["normalization"]["enable"] = False
    torch.manual_seed(1856)
    got = tn.generate_points_batched(cfg, 5, 2, (96, 3), pin=False)
    assert np.array_equal(got.numpy(), golden_cpu["gp_plain"])


def test_trimesh_chamfer_restatement_matches_reference_kdtree(oracle, golden_cpu):
    """oracle.trimesh_chamfer (brute force) vs the reference's KD-tree compute_trimesh_chamfer outputs."""
    pcs = golden_cpu["tm_pcs"]
    for j in range(5):
        for k in range(5):
            assert oracle.trimesh_chamfer(pcs[j], pcs[k]) == pytest.approx(float(golden_cpu["tm_cd"][j, k]), rel=1e-9, abs=1e-15)


def test_oracle_reproduces_reference_full_model_loop(hp, oracle, golden_cpu):
    """The per-sample loop of the reference's FullModel.forward (model/full_model.py:67-74), run on CPU by the golden script:
    re-drawing the input clouds in the same order from the same seed and applying the oracle MLP to the recorded hypernetwork
    output gives the recorded reconstruction (host-side sampling order + weight layout pinned without a GPU)."""
    import torch

    g = golden_cpu
    B, N, epoch, seed = (int(v) for v in g["fm_meta"])
    pcfg = {"target_network_input": {"constant": False, "normalization": {"enable": True, "type": "progressive", "epoch": 100}}}
    torch.manual_seed(seed)
    pts = hp.generate_points_batched(pcfg, epoch, B, (N, 3), pin=False)
    y = oracle.target_network_forward(g["fm_weights"], pts.numpy(), [32, 64, 128, 64], True)
    err = np.abs(np.transpose(y, (0, 2, 1)) - g["fm_rec"]).max() / np.abs(g["fm_rec"]).max()
    assert err < 1e-6, err


def test_jsd_restatement_matches_reference(oracle, golden_cpu):
    """Occupancy-grid JSD (utils/metrics.py:243-359) restated with an exhaustive float64 nearest-centre search against the values
    the reference's scikit-learn implementation produced."""
    g = golden_cpu
    assert np.array_equal(oracle.unit_cube_grid(8, True), g["jsd_grid8"])
    for res in (8, 28):
        assert np.array_equal(oracle.occupancy_counts(g["jsd_smp"], res), g[f"jsd_cnt_r{res}"].astype(np.float64))
        got = oracle.jsd_between_point_cloud_sets(g["jsd_smp"], g["jsd_ref"], res)
        assert abs(got - float(g[f"jsd_r{res}"])) <= 1e-9 * abs(float(g[f"jsd_r{res}"]))


def test_emd_exhausted_points_premise_of_the_compaction(oracle):
    """The premise of emd_compact_kernel, checked on the CPU restatement of approxmatch.cu: the clamp of
    approxmatch.cu:140 leaves most points of the second cloud with remainR == 0 EXACTLY (not merely small), so their later terms
    are exact zeros; and a point whose remainR is 0 receives nothing more (its column of `match` stops growing), which is what
    allows the kernels to leave it out."""
    rng = np.random.default_rng(11)
    n = 512
    a = (rng.random((2, n, 3), dtype=np.float32) - 0.5)
    b = (rng.random((2, n, 3), dtype=np.float32) - 0.5)
    match, temp = oracle.approx_match(a, b)
    remain_l, remain_r = temp[:, :n], temp[:, n:2 * n]
    assert (remain_r >= 0).all() and (remain_l >= 0).all()
    assert (remain_r == 0).mean() > 0.9            # exhausted exactly, the common case
    received = match.sum(axis=2)                    # mass point l of the second cloud received, match is [B, m, n]
    assert (received[remain_r == 0] > 0.99).all()   # exhausted points are full (multiR = 1 for n == m) ...
    assert (received <= 1.0 + 1e-4).all()           # ... and nobody is over-full
