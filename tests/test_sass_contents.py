"""The shipped library is sm_100a machine code of the kinds DESIGN.md claims (no GPU needed: cuobjdump reads the cubin).
  * Chamfer / EMD / metrics: packed fp32 (FFMA2 / FADD2), bulk-TMA staging (UBLKCP), MUFU.EX2 -- K = 3 stays on the CUDA cores;
  * TargetNetwork forward: tcgen05 (UTCHMMA) with tensor-memory loads / stores (LDTM / STTM);
  * TargetNetwork backward: legacy tensor path (HMMA.1688.F32.TF32) with the running gradient in tensor memory;
  * programmatic dependent launches: griddepcontrol.launch_dependents / .wait (PREEXIT / ACQBULK) in the Chamfer ring -> tail / unpack
    pair and along the EMD auction's chain of kernels, incl. its compaction kernel."""
import collections
import re
import shutil
import subprocess

import pytest


def test_library_sass_matches_the_design(hp):
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([cuobjdump, "-sass", hp._native.LIB_PATH], capture_output=True, text=True, timeout=600).stdout
    except (OSError, subprocess.TimeoutExpired):
        pytest.skip("cuobjdump not available")
    assert "arch = sm_100a" in out and "arch = sm_90" not in out and "arch = sm_80" not in out
    ops = collections.Counter(m.group(1) for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", out, re.M))
    fam = collections.Counter()
    for op, n in ops.items():
        fam[op.split(".")[0]] += n
    assert fam["FFMA2"] > 1000 and fam["UBLKCP"] > 10 and ops["MUFU.EX2"] > 100            # Chamfer / EMD on the CUDA cores
    assert fam["UTCHMMA"] >= 84 and fam["LDTM"] > 0 and fam["STTM"] > 0                    # tcgen05 forward, tensor-memory traffic
    assert ops["HMMA.1688.F32.TF32"] > 2000                                                # mma.sync backward (and mode-2 forward)
    functions = set(re.findall(r"Function : (\S+)", out))
    for needle in ("nn_ring_kernel", "nn_ring_tail_kernel", "tn_tc5_forward_kernel", "tn_mma_backward_kernel", "tn_mma_forward_kernel",
                   "tn_forward_kernel", "tn_backward_kernel", "emd_pass_kernel", "pairwise_cd_kernel", "batch_pairwise_dist_kernel"):
        assert any(needle in f for f in functions), needle
    assert any("emd_compact_kernel" in f for f in functions)
    # per function: the opcodes between its "Function :" line and the next
    per = {}
    for chunk in re.split(r"^\s*Function : ", out, flags=re.M)[1:]:
        name, _, body = chunk.partition("\n")
        per[name.strip()] = set(re.findall(r"\b(PREEXIT|ACQBULK)\b", body))
    def has(needle, op):
        hits = [ops_ for f, ops_ in per.items() if needle in f]
        return bool(hits) and all(op in ops_ for ops_ in hits)
    assert has("nn_ring_kernel", "PREEXIT")                                                 # lets the tail / unpack kernel in early
    assert has("nn_ring_unpack_kernel", "ACQBULK") and has("nn_ring_tail_kernel", "ACQBULK")
    for k in ("emd_pass_kernel", "emd_combine_kernel", "emd_compact_kernel", "emd_init_kernel", "emd_cost_finish_kernel"):
        assert has(k, "PREEXIT") and has(k, "ACQBULK"), k                                   # every link of the auction's chain
