"""Every kernel and every stream encoding on small inputs, self-checked (compact stream vs
plain CSR walk, export round trip, batch vs loop): tools/kernel_coverage.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_kernel_coverage_script():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kernel_coverage.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "KERNEL_COVERAGE_OK" in out.stdout
    # the inputs are built to reach the dense, per-cell and tile encodings
    lines = [l for l in out.stdout.splitlines() if "enc=" in l]
    assert any("'D16': 0" not in l for l in lines)
    assert any("'S32': 0" not in l for l in lines)
    assert any("'S16': 0" not in l for l in lines)
    assert all("'T': 0" not in l for l in lines)
