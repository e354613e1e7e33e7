"""CPU-side checks of the C ABI's host logic (no kernel is launched, no GPU needed): size helpers, argument
validation and error reporting of libhp_b200.so."""
import ctypes

import pytest


@pytest.fixture(scope="module")
def lib(hp):
    return hp._native.load()


def _dims(*d):
    return (ctypes.c_int * len(d))(*d), len(d) - 1


def test_target_network_num_weights(lib):
    cd, nl = _dims(3, 32, 64, 128, 64, 3)
    assert lib.hp_target_network_num_weights(nl, cd, 1) == 19011  # settings/*.json.sample layer_out_channels
    assert lib.hp_target_network_num_weights(nl, cd, 0) == 18720
    cd2, nl2 = _dims(3, 16, 8, 3)
    assert lib.hp_target_network_num_weights(nl2, cd2, 0) == 3 * 16 + 16 * 8 + 8 * 3
    assert lib.hp_target_network_num_weights(0, cd, 1) == -1
    bad, nlb = _dims(3, 0, 3)
    assert lib.hp_target_network_num_weights(nlb, bad, 1) == -1


def test_workspace_sizes_are_monotone_and_cover_key_arrays(lib):
    # ring forward: two 64-bit key arrays of b*(n+m) entries plus small bookkeeping
    b, n, m = 32, 2048, 2048
    ws = lib.hp_chamfer_workspace_bytes(b, n, m)
    assert ws >= 8 * b * (n + m) and ws < 8 * b * (n + m) + (1 << 16)
    assert lib.hp_chamfer_workspace_bytes(2 * b, n, m) > ws
    assert lib.hp_chamfer_workspace_bytes(0, n, m) == 16
    assert lib.hp_chamfer_inverse_ints(b, n, m, 1) == b * (n + 2 * m)
    assert lib.hp_chamfer_inverse_ints(b, 1000, 500, 2) == b * (500 + 2 * 1000)
    assert lib.hp_chamfer_inverse_ints(b, 40000, m, 1) == 0  # too large for the in-kernel inverse
    assert lib.hp_approxmatch_workspace_bytes(b, n, m) >= 4 * b * 9 * (n + m)
    assert lib.hp_emd_cost_workspace_bytes(4096, n, m) >= 4 * 4096 * 2 * (n + m)
    cd, nl = _dims(3, 32, 64, 128, 64, 3)
    assert lib.hp_target_network_backward_workspace_bytes(64, 2048, nl, cd, 1) >= 4 * 64
    assert lib.hp_target_network_backward_workspace_bytes(0, 2048, nl, cd, 1) == 16


def test_argument_validation_reports_errors_without_touching_the_gpu(hp, lib):
    INVALID = hp._native.HP_ERR_INVALID_ARGUMENT
    assert lib.hp_nndistance(-1, 4, None, 4, None, None, None, None, None, None) == INVALID
    assert b"negative size" in lib.hp_last_error_message()
    assert lib.hp_nndistance(2, 4, None, 4, None, None, None, None, None, None) == INVALID  # null pointers
    assert lib.hp_nndistance(2, 0, None, 4, None, None, None, None, None, None) == INVALID  # one empty set
    assert lib.hp_nndistance(0, 4, None, 4, None, None, None, None, None, None) == hp._native.HP_OK  # empty batch: no-op
    assert lib.hp_pairwise_cd(4, 4, 16, 16, None, None, 3, 2, None, None) == INVALID  # bad row range
    assert lib.hp_emd_cost_pairs(-1, 8, 8, None, None, None, None, None, None, 0, None) == INVALID
    cd, nl = _dims(3, 32, 64, 128, 64, 4)  # must map 3 -> 3
    assert lib.hp_target_network_forward(1, 8, nl, cd, 1, None, None, 24, None, 0, None) == INVALID
    assert b"3 -> 3" in lib.hp_last_error_message()
    with pytest.raises(RuntimeError, match="invalid argument"):
        hp._native.check(INVALID, "demo")
    assert lib.hp_error_string(hp._native.HP_ERR_WORKSPACE) == b"workspace too small"
    assert lib.hp_version() == 100


def test_target_network_mode_switch_validates(hp, lib):
    INVALID = hp._native.HP_ERR_INVALID_ARGUMENT
    assert lib.hp_target_network_set_mode(7) == INVALID
    assert b"neither 0" in lib.hp_last_error_message()
    assert lib.hp_target_network_set_mode(1) == hp._native.HP_OK   # fp32 FFMA kernels
    assert lib.hp_target_network_set_mode(2) == hp._native.HP_OK   # 3xTF32 on mma.sync only
    assert lib.hp_target_network_set_mode(0) == hp._native.HP_OK   # back to the default (3xTF32: tcgen05 forward, mma.sync backward)
    with pytest.raises(ValueError):
        hp.target_network_set_mode("bf16")


def test_host_io_split_boundaries(hp):
    """ChamferStepGraph._host_io_bounds: which batch slices run_from_host_loss_only runs one after the other (host logic only)."""
    G = hp.graphs.ChamferStepGraph
    obj = object.__new__(G)
    bounds = lambda split, batch=5: G._host_io_bounds(obj, batch, 700, 900, split)  # noqa: E731
    assert bounds(None) == bounds(False) == bounds(0) == bounds(1) == [0, 5]        # the default: one part
    assert bounds(True) == bounds(2) == [0, 2, 5]
    assert bounds(3) == [0, 1, 3, 5] and bounds(9) == [0, 1, 2, 3, 4, 5]            # never more parts than clouds
    assert bounds([0, 4, 5]) == [0, 4, 5] and bounds((0, 5)) == [0, 5]
    for bad in ([1, 5], [0, 3, 3, 5], [0, 4]):
        with pytest.raises(ValueError):
            bounds(bad)
    assert bounds(2, batch=8) == [0, 4, 8]
    # a part the fused step cannot take (clouds above 4096 points) falls back to one part
    assert G._host_io_bounds(obj, 4, 5000, 5000, 2) == [0, 4]
