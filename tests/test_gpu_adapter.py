"""GPU: the C++ drop-in adapter (include/orbmatch_b200/ORBmatcher.hpp, the reference's call signatures)
against the reference's own ORBmatcher.cc on the same C++ objects.  The binary is built where the
reference is mounted (`make -C oracle adapter`) and travels to the GPU box under oracle/_ref/."""
import os
import subprocess

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


def test_cpp_adapter_matches_reference():
    exe = os.path.join(ROOT, "oracle", "_ref", "adapter_parity")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/adapter_parity not built (needs /root/reference at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "ADAPTER PARITY OK" in r.stdout
