"""CPU: host-side logic of the sharded metrics (row-block partition, (min,argmin) merges, 1-NNA from
blocks) against the oracle / golden vectors, single process and world_size=2 over gloo."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_mmd_cov_and_knn_match_reference_golden(hp, golden_cpu):
    g = golden_cpu
    r = hp.metrics.mmd_cov(torch.from_numpy(g["mc_M"]))
    assert float(r["mmd(Fidelity)"]) == pytest.approx(float(g["mc_mmd"]), rel=1e-6)
    assert float(r["cov(Coverage)"]) == pytest.approx(float(g["mc_cov"]), rel=1e-6)
    assert float(r["mmd_smp"]) == pytest.approx(float(g["mc_mmd_smp"]), rel=1e-6)
    k = hp.metrics.knn(torch.from_numpy(g["knn_Mxx"]), torch.from_numpy(g["knn_Mxy"]), torch.from_numpy(g["knn_Myy"]), 1)
    for key in ("tp", "fp", "fn", "tn", "precision", "recall", "acc_t", "acc_f", "acc"):
        assert float(k[key]) == pytest.approx(float(g["knn_" + key]), rel=1e-6), key


def test_knn_and_mmd_cov_match_oracle_on_random_matrices_with_ties(hp, oracle):
    gen = torch.Generator().manual_seed(4)
    for n0, n1 in [(5, 7), (16, 16), (1, 3)]:
        Mxx = torch.randint(0, 6, (n0, n0), generator=gen).float()
        Myy = torch.randint(0, 6, (n1, n1), generator=gen).float()
        Mxy = torch.randint(0, 6, (n0, n1), generator=gen).float()
        k = hp.metrics.knn(Mxx, Mxy, Myy, 1)
        ok = oracle.knn(Mxx.numpy(), Mxy.numpy(), Myy.numpy(), 1)
        for key in ("tp", "fp", "fn", "tn", "acc"):
            assert float(k[key]) == pytest.approx(ok[key], rel=1e-6), (n0, n1, key)
        r = hp.metrics.mmd_cov(Mxy.t().contiguous())
        o = oracle.mmd_cov(Mxy.t().numpy())
        for key in o:
            assert float(r[key]) == pytest.approx(float(o[key]), rel=1e-6)


def test_shard_rows_partition(hp):
    for n in (0, 1, 7, 1000):
        for world in (1, 2, 3, 8):
            blocks = [hp.metrics.shard_rows(n, r, world) for r in range(world)]
            covered = [i for b, e in blocks for i in range(b, e)]
            assert covered == list(range(n))
            assert max(e - b for b, e in blocks) <= (n + world - 1) // world


def test_upper_triangle_pairs_and_nearest_other(hp):
    """Closed-form pair list of the strict upper triangle; nearest-other vector from any chunking of it equals the
    column minima of the symmetric matrix with an infinite diagonal."""
    M = hp.metrics
    for n in (0, 1, 2, 3, 7, 100, 1000):
        tot = n * (n - 1) // 2
        r, s_ = M.upper_triangle_pairs(n, 0, tot, "cpu")
        tr = torch.triu_indices(n, n, 1)
        assert r.dtype == torch.int32 and torch.equal(r.long(), tr[0]) and torch.equal(s_.long(), tr[1]), n
        for world in (1, 2, 3, 8):
            chunks = [M.shard_pairs(tot, k, world) for k in range(world)]
            assert [i for b, e in chunks for i in range(b, e)] == list(range(tot))
            assert max(e - b for b, e in chunks) <= (tot + world - 1) // world
    r, s_ = M.upper_triangle_pairs(200000, 200000 * 199999 // 2 - 2, 200000 * 199999 // 2, "cpu")  # far beyond float32 precision
    assert r.tolist() == [199997, 199998] and s_.tolist() == [199999, 199999]
    gen = torch.Generator().manual_seed(3)
    for n in (2, 9, 40):
        A = torch.randint(0, 4, (n, n), generator=gen).float()  # small integers: ties
        A = torch.maximum(A, A.t())
        want = (A + torch.diag(torch.full((n,), float("inf")))).min(0).values
        tot = n * (n - 1) // 2
        parts = []
        for k in range(3):
            p0, p1 = M.shard_pairs(tot, k, 3)
            r, s_ = M.upper_triangle_pairs(n, p0, p1, "cpu")
            parts.append(M.nearest_other_from_pair_values(n, r, s_, A[r.long(), s_.long()]))
        assert torch.equal(torch.stack(parts).min(0).values, want)
    assert M.nearest_other_from_pair_values(1, *M.upper_triangle_pairs(1, 0, 0, "cpu"), torch.empty(0)).tolist() == [float("inf")]


def _worker(rank, world, port, path, n_ref, n_smp):
    sys.path.insert(0, REPO)
    import importlib

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
    M = hp.metrics
    d = torch.load(path)
    rb, re_ = M.shard_rows(n_ref, rank, world)
    sb, se = M.shard_rows(n_smp, rank, world)
    res = M.mmd_cov_from_block(d["M_rs"][rb:re_].contiguous(), rb, n_ref, None)
    one = M.knn_from_blocks(d["M_rr"][rb:re_].contiguous(), d["M_rs"][rb:re_].contiguous(), d["M_ss"][sb:se].contiguous(),
                            rb, sb, n_ref, n_smp, 1, False, None)
    out = {k: float(v) for k, v in res.items()}
    out.update({"knn_" + k: float(v) for k, v in one.items()})
    # the symmetric formulation used by compute_all_metrics: pair lists of the upper triangles instead of row blocks
    near = []
    for key, n in (("S_rr", n_ref), ("S_ss", n_smp)):
        p0, p1 = M.shard_pairs(n * (n - 1) // 2, rank, world)
        r, s_ = M.upper_triangle_pairs(n, p0, p1, "cpu")
        near.append(M.nearest_other_from_pair_values(n, r, s_, d[key][r.long(), s_.long()], None))
    xy_row, _ = M.row_min_gathered(d["M_rs"][rb:re_].contiguous(), n_ref, None)
    xy_col, _ = M.col_min_merged(d["M_rs"][rb:re_].contiguous(), rb, n_ref, None)
    out.update({"sym_" + k: float(v) for k, v in M.knn1_from_nearest(near[0], xy_row, xy_col, near[1]).items()})
    torch.save(out, f"{path}.rank{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_ref,n_smp", [(9, 6), (4, 4)])
def test_world_size_2_gloo_equals_single_process(hp, tmp_path, n_ref, n_smp):
    gen = torch.Generator().manual_seed(n_ref)
    # small integers -> plenty of ties, so the lowest-index merge rule is exercised across ranks
    d = {"M_rs": torch.randint(0, 5, (n_ref, n_smp), generator=gen).float(),
         "M_rr": torch.randint(0, 5, (n_ref, n_ref), generator=gen).float(),
         "M_ss": torch.randint(0, 5, (n_smp, n_smp), generator=gen).float()}
    d["S_rr"], d["S_ss"] = torch.maximum(d["M_rr"], d["M_rr"].t()), torch.maximum(d["M_ss"], d["M_ss"].t())  # symmetric
    path = str(tmp_path / "mats.pt")
    torch.save(d, path)
    single = {k: float(v) for k, v in hp.metrics.mmd_cov_from_block(d["M_rs"], 0, n_ref, None).items()}
    single.update({"knn_" + k: float(v) for k, v in hp.metrics.knn(d["M_rr"], d["M_rs"], d["M_ss"], 1).items()})
    single.update({"sym_" + k: float(v) for k, v in hp.metrics.knn(d["S_rr"], d["M_rs"], d["S_ss"], 1).items()})
    mp.spawn(_worker, args=(2, _free_port(), path, n_ref, n_smp), nprocs=2, join=True)
    for rank in range(2):
        got = torch.load(f"{path}.rank{rank}")
        assert got == single, (rank, got, single)


def test_knn_general_k_matches_reference_golden_semantics():
    """knn(k>1) votes over the k nearest columns like utils/metrics.py:162-191; checked against a direct numpy vote."""
    import importlib

    import numpy as np
    import torch

    m = importlib.import_module("3d-point-clouds-autocomplete_b200.metrics")
    g = torch.Generator().manual_seed(4)
    n0, n1, k = 9, 7, 3
    Mxx = torch.rand(n0, n0, generator=g)
    Mxx = (Mxx + Mxx.t()) / 2
    Myy = torch.rand(n1, n1, generator=g)
    Myy = (Myy + Myy.t()) / 2
    Mxy = torch.rand(n0, n1, generator=g)
    s = m.knn(Mxx, Mxy, Myy, k)
    M = np.block([[Mxx.numpy(), Mxy.numpy()], [Mxy.numpy().T, Myy.numpy()]]) + np.diag(np.full(n0 + n1, np.inf))
    label = np.r_[np.ones(n0), np.zeros(n1)]
    pred = np.array([(label[np.argsort(M[:, c], kind="stable")[:k]].sum() >= k / 2) for c in range(n0 + n1)], float)
    assert float(s["acc"]) == np.mean(pred == label).astype(np.float32)
    assert float(s["tp"]) == float((pred * label).sum()) and float(s["tn"]) == float(((1 - pred) * (1 - label)).sum())
    # k = 1 path agrees with the general path
    s1 = m.knn(Mxx, Mxy, Myy, 1)
    M1 = np.array([(label[np.argmin(M[:, c])]) for c in range(n0 + n1)], float)
    assert float(s1["acc"]) == np.mean(M1 == label).astype(np.float32)


def test_jsd_host_pieces_match_reference_golden(hp, golden_cpu):
    """Grid construction and the JSD formula (utils/metrics.py:243-262,323-340) against values produced by the reference's own
    functions; the nearest-grid-centre search itself (NN kernel) is checked on the GPU in tests/test_metrics_gpu.py."""
    g = golden_cpu
    M = hp.metrics
    grid, spacing = M.unit_cube_grid_point_cloud(8, True)
    assert grid.dtype == np.float32 and np.array_equal(grid, g["jsd_grid8"]) and spacing == float(g["jsd_spacing8"])
    assert M.unit_cube_grid_point_cloud(5, False)[0].shape == (5, 5, 5, 3)
    for res in (8, 28):
        cells = M.unit_cube_grid_point_cloud(res, True)[0].astype(np.float64)
        def counts(pcs):  # brute-force float64 nearest centre
            pts = pcs.reshape(-1, 3).astype(np.float64)
            idx = ((pts[:, None, :] - cells[None, :, :]) ** 2).sum(-1).argmin(1)
            return np.bincount(idx, minlength=len(cells)).astype(np.float64)
        c_smp, c_ref = counts(g["jsd_smp"]), counts(g["jsd_ref"])
        assert np.array_equal(c_smp, g[f"jsd_cnt_r{res}"].astype(np.float64))
        assert M.jensen_shannon_divergence(c_smp, c_ref) == pytest.approx(float(g[f"jsd_r{res}"]), rel=1e-9)
    with pytest.raises(ValueError):
        M.jensen_shannon_divergence(np.array([1.0, -1.0]), np.array([1.0, 1.0]))
