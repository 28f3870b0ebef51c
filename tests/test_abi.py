"""The C-ABI library builds, loads and exports every symbol include/ssak_b200.h declares (no
compute calls: there is no GPU on the CPU box)."""
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ssak_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"SSAK_API\s+[\w\s\*]+?\b(ssak_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = _header_symbols()
    for name in ("ssak_ctc_loss_forward", "ssak_ctc_loss_backward", "ssak_ctc_logits_forward",
                 "ssak_ctc_logits_backward", "ssak_forced_align", "ssak_ctc_greedy",
                 "ssak_ctc_loss_host", "ssak_forced_align_host", "ssak_ctc_greedy_host", "ssak_b200_version"):
        assert name in syms


def test_library_builds_and_exports_header_symbols():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "ssak_b200"))
    from ssak_b200 import build as B
    lib_path = B.build()
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in _header_symbols() if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    stray = [s for s in exported if s.startswith("ssak_") and s not in _header_symbols()]
    assert not stray, f"exported but not declared: {stray}"


def test_ctypes_table_matches_header_and_loads():
    import ssak_b200
    from ssak_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _header_symbols()
    L = ssak_b200.lib()
    assert L.ssak_b200_version() >= 100
    assert L.ssak_b200_strerror(0) == b"ok" and b"workspace" in L.ssak_b200_strerror(-3)
    # pure host-side entry points: sizes grow with the problem, unsupported shapes report 0
    small = L.ssak_ctc_loss_workspace_bytes(100, 4, 20, 1)
    assert 0 < small < L.ssak_ctc_loss_workspace_bytes(200, 4, 20, 1)
    assert L.ssak_ctc_loss_workspace_bytes(100, 4, 20, 0) < small
    assert L.ssak_ctc_loss_workspace_bytes(100, 4, 100000, 1) == 0
    assert L.ssak_ctc_loss_workspace_bytes(300001, 1, 10, 1) == 0      # T limit of the loss (exact fp32 offsets)
    assert L.ssak_ctc_loss_workspace_bytes(300000, 1, 10, 0) > 0
    assert L.ssak_align_workspace_bytes(16, 30000, 8000) > 16 * 30000 * 8001 // 4
    assert L.ssak_align_workspace_bytes(1, 10, 100000) == 0


def test_sass_uses_bulk_copy_and_mufu():
    """The kernels are Blackwell-native where the design says so: cp.async.bulk (UBLKCP) + mbarrier
    (SYNCS) emission prefetch, MUFU ex2/lg2 recursions."""
    from ssak_b200 import build as B
    try:
        sass = subprocess.check_output(["cuobjdump", "-sass", B.build()], text=True)
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("cuobjdump not available")
    for mnemonic in ("UBLKCP", "SYNCS", "MUFU.EX2", "MUFU.LG2"):
        assert mnemonic in sass, mnemonic


def test_no_oracle_or_cpu_fallback_in_product():
    """The product package never imports the oracle and has no CPU path."""
    pkg = os.path.join(ROOT, "ssak_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in txt and "from oracle" not in txt and "oracle." not in txt, fn
