"""GPU parity tests for the set-vs-set metrics path (pairwise CD/EMD matrices, MMD/COV/1-NNA)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sets(na, nb, n, m, seed):
    g = torch.Generator().manual_seed(seed)
    a = (torch.rand(na, n, 3, generator=g) - 0.5).contiguous()
    b = (torch.rand(nb, m, 3, generator=g) - 0.5).contiguous()
    return a, b


@pytest.mark.parametrize("na,nb,n,m", [(3, 4, 256, 256), (2, 3, 300, 129), (2, 2, 2048, 2048), (1, 2, 5000, 2500), (2, 2, 1, 7), (1, 1, 4100, 33)])
def test_pairwise_cd_vs_oracle(hp, oracle, na, nb, n, m):
    a, b = _sets(na, nb, n, m, seed=n + m)
    cd = hp.pairwise_cd(a.to(DEV), b.to(DEV))
    ocd = oracle.pairwise_cd_direct(a.numpy(), b.numpy())
    # every min is bit-exact (same arithmetic as nn_distance); only the order of the two mean reductions differs
    np.testing.assert_allclose(cd.cpu().numpy(), ocd, rtol=2e-6)
    # from the nearest-neighbour kernel of this library: identical mins, so very tight
    ad, bd = a.to(DEV), b.to(DEV)
    for r in range(na):
        d1, _, d2, _ = hp.NNDistance(ad[r:r + 1].expand(nb, -1, -1).contiguous(), bd)
        ref = d1.double().mean(1) + d2.double().mean(1)
        torch.testing.assert_close(cd[r].double(), ref, rtol=2e-6, atol=0)


def test_pairwise_cd_row_blocks_are_bit_identical(hp):
    a, b = _sets(7, 5, 512, 512, seed=2)
    ad, bd = a.to(DEV), b.to(DEV)
    full = hp.pairwise_cd(ad, bd)
    parts = torch.cat([hp.pairwise_cd(ad, bd, 0, 3), hp.pairwise_cd(ad, bd, 3, 4), hp.pairwise_cd(ad, bd, 4, 7)])
    assert torch.equal(full, parts)
    assert torch.equal(full, hp.pairwise_cd(ad, bd))
    # symmetry of the definition: cd(a_r, b_s) == cd(b_s, a_r) up to the order of the two means
    swapped = hp.pairwise_cd(bd, ad)
    torch.testing.assert_close(full, swapped.t(), rtol=1e-6, atol=0)


def test_pairwise_cd_pair_list_and_symmetric_nearest_other(hp):
    """hp_pairwise_cd_pairs gives the bits of the matrix kernel for the same cloud pairs; nearest_other_cd (strict upper
    triangle only) equals the column minima of the full self-distance matrix up to the order of the two means."""
    a, b = _sets(9, 6, 300, 257, seed=12)
    ad, bd = a.to(DEV), b.to(DEV)
    full = hp.pairwise_cd(ad, bd)
    g = torch.Generator().manual_seed(1)
    r = torch.randint(0, 9, (40,), generator=g, dtype=torch.int32).to(DEV)
    s_ = torch.randint(0, 6, (40,), generator=g, dtype=torch.int32).to(DEV)
    got = hp.metrics.pairwise_cd_pairs(ad, bd, r, s_)
    assert torch.equal(got, full[r.long(), s_.long()])
    assert hp.metrics.pairwise_cd_pairs(ad, bd, r[:0], s_[:0]).numel() == 0
    with pytest.raises(RuntimeError):
        hp.metrics.pairwise_cd_pairs(ad, bd, r, s_ + 6)
    self_full = hp.pairwise_cd(ad, ad)
    want = (self_full + torch.diag(torch.full((9,), float("inf"), device=DEV))).min(0).values
    near = hp.metrics.nearest_other_cd(ad)
    torch.testing.assert_close(near, want, rtol=1e-6, atol=0)
    upper = torch.triu_indices(9, 9, 1).to(DEV)
    assert torch.equal(hp.metrics.pairwise_cd_pairs(ad, ad, upper[0].int(), upper[1].int()), self_full[upper[0], upper[1]])
    one = hp.metrics.nearest_other_cd(ad[:1])
    assert one.shape == (1,) and torch.isinf(one).all()


def test_pairwise_cd_vs_reference_expansion_form(hp, oracle):
    """vs the reference's own dist_chamfer arithmetic (expansion form, torch port): 1e-5 relative."""
    a, b = _sets(3, 3, 1024, 1024, seed=5)
    cd = hp.pairwise_cd(a.to(DEV), b.to(DEV)).cpu().numpy()
    ref = oracle.pairwise_cd_expansion(a.numpy(), b.numpy(), batch_size=2)
    np.testing.assert_allclose(cd, ref, rtol=1e-5)


def test_pairwise_emd_vs_oracle_and_reference_extension(hp, oracle, ref_ext):
    """The match-free fused cost (hp_emd_cost_pairs, several batched calls) against the C oracle's approxmatch + matchcost and
    against the reference's own extension composed like emd_approx (utils/metrics.py:71-76): ApproxMatch -> MatchCost -> / N."""
    a, b = _sets(3, 4, 256, 256, seed=6)
    ad, bd = a.to(DEV), b.to(DEV)
    emd = hp.pairwise_emd(ad, bd, max_pairs_per_call=5)  # forces several batched calls
    np.testing.assert_allclose(emd.cpu().numpy(), oracle.pairwise_emd(a.numpy(), b.numpy()), rtol=1e-5)
    for r in range(3):
        rep = ad[r:r + 1].expand(4, -1, -1).contiguous()
        match, _temp = ref_ext.ApproxMatch(rep, bd)
        row = ref_ext.MatchCost(rep, bd, match) / 256.0
        torch.testing.assert_close(emd[r], row, rtol=1e-5, atol=1e-8)


def test_compute_all_metrics_vs_oracle(hp, oracle):
    smp, ref = _sets(6, 5, 128, 128, seed=9)
    res = hp.compute_all_metrics(smp.to(DEV), ref.to(DEV), batch_size=3, chamfer_loss=hp.ChamferLoss(), one_nn=True)
    ores = oracle.compute_all_metrics(smp.numpy(), ref.numpy(), with_emd=True, with_1nn=True)
    assert set(res) == set(ores), (sorted(res), sorted(ores))
    for k in ("mmd(Fidelity)-CD", "cov(Coverage)-CD", "mmd_smp-CD", "mmd(Fidelity)-EMD", "cov(Coverage)-EMD", "mmd_smp-EMD"):
        assert k in res
    for k, v in ores.items():
        assert v == pytest.approx(float(res[k]), rel=5e-5), k
        assert res[k].dim() == 0 and res[k].is_cuda


def test_compute_all_metrics_reference_keys_without_1nn(hp):
    smp, ref = _sets(4, 4, 64, 64, seed=1)
    res = hp.compute_all_metrics(smp.to(DEV), ref.to(DEV), 2, hp.ChamferLoss())
    assert sorted(res) == sorted(["mmd(Fidelity)-CD", "cov(Coverage)-CD", "mmd_smp-CD",
                                  "mmd(Fidelity)-EMD", "cov(Coverage)-EMD", "mmd_smp-EMD"])


def test_dist_chamfer_and_emd_approx(hp, oracle, ref_ext):
    """dist_chamfer against the reference's arithmetic (expansion-form matrix of champfer_loss.py:19-35, torch port pinned to the
    reference by tests/test_oracle_golden.py, then P.min(1) / P.min(2) like utils/metrics.py:78-83); emd_approx against the
    reference extension's ApproxMatch + MatchCost."""
    a, b = _sets(3, 3, 200, 150, seed=3)
    ad, bd = a.to(DEV), b.to(DEV)
    dl, dr = hp.metrics.dist_chamfer(ad, bd, hp.ChamferLoss())
    P = oracle.batch_pairwise_dist_torch(a, b)
    torch.testing.assert_close(dl.cpu(), P.min(1)[0], rtol=0, atol=2e-6)
    torch.testing.assert_close(dr.cpu(), P.min(2)[0], rtol=0, atol=2e-6)
    a2, b2 = _sets(2, 2, 128, 128, seed=4)
    e = hp.metrics.emd_approx(a2.to(DEV), b2.to(DEV))
    match, _temp = ref_ext.ApproxMatch(a2.to(DEV), b2.to(DEV))
    torch.testing.assert_close(e, ref_ext.MatchCost(a2.to(DEV), b2.to(DEV), match) / 128.0, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(e.cpu().numpy(), oracle.match_cost(a2.numpy(), b2.numpy()) / 128.0, rtol=1e-5)


def test_jsd_vs_oracle_on_seeded_sets(hp, oracle):
    """Fresh seeded sets (not the golden inputs): the NN-kernel occupancy counts equal the oracle's exhaustive float64 search."""
    g = torch.Generator().manual_seed(21)
    smp = ((torch.rand(7, 300, 3, generator=g) - 0.5) * 0.57).numpy()
    ref = ((torch.rand(4, 500, 3, generator=g) - 0.5) * 0.5).numpy()
    for res in (5, 16):
        cnt = hp.metrics.entropy_of_occupancy_grid(smp, res, True)[1]
        ocnt = oracle.occupancy_counts(smp, res)
        assert np.abs(cnt - ocnt).sum() <= 2, res   # a point within fp32 rounding of a cell boundary may fall either way
        got = hp.metrics.jsd_between_point_cloud_sets(smp, ref, res)
        assert got == pytest.approx(oracle.jsd_between_point_cloud_sets(smp, ref, res), rel=1e-4, abs=1e-7)


def test_jsd_vs_reference_golden(hp, golden_cpu):
    """jsd_between_point_cloud_sets / entropy_of_occupancy_grid (utils/metrics.py:265-320) with the nearest grid centre of every
    point found by the NN ring kernel, against the reference's scikit-learn KD-tree implementation (golden vectors)."""
    g = golden_cpu
    for res in (8, 28):
        ent, cnt = hp.metrics.entropy_of_occupancy_grid(g["jsd_smp"], res, True)
        assert np.array_equal(cnt, g[f"jsd_cnt_r{res}"].astype(np.float64)), res
        assert ent == pytest.approx(float(g[f"jsd_ent_r{res}"]), rel=1e-9)
        jsd = hp.metrics.jsd_between_point_cloud_sets(g["jsd_smp"], torch.from_numpy(g["jsd_ref"]).to(DEV), res)
        assert jsd == pytest.approx(float(g[f"jsd_r{res}"]), rel=1e-9)
    # C5-sized set: 1000 clouds x 2048 points against the 28^3 grid in one launch; identical sets have zero divergence
    big = (torch.rand(1000, 2048, 3, generator=torch.Generator().manual_seed(0)) - 0.5).to(DEV) * 0.55
    ent, cnt = hp.metrics.entropy_of_occupancy_grid(big, 28, True)
    assert cnt.sum() == 1000 * 2048 and 0.0 < ent < 1.0
    assert abs(hp.metrics.jsd_between_point_cloud_sets(big[:500], big[:500])) < 1e-12
    assert hp.metrics.jsd_between_point_cloud_sets(big[:500], big[500:] * 0.5) > 0.1
