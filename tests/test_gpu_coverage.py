"""Every kernel and every stream encoding on small inputs, self-checked (compact stream vs
plain CSR walk, export round trip, batch vs loop): tools/kernel_coverage.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env):
    env = dict(os.environ, **extra_env)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kernel_coverage.py")],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "KERNEL_COVERAGE_OK" in out.stdout
    return [l for l in out.stdout.splitlines() if "enc=" in l]


@pytest.mark.gpu
def test_kernel_coverage_script():
    lines = _run({})
    # the inputs are built to reach the dense, per-cell and tile encodings
    assert any("'D16': 0" not in l for l in lines)
    assert any("'S32': 0" not in l for l in lines)
    assert any("'S16': 0" not in l for l in lines)
    assert all("'T': 0" not in l for l in lines)


@pytest.mark.gpu
def test_kernel_coverage_under_the_checked_build():
    """-DTB_CHECK=1: every descriptor, ring, gather and decode index of the hot kernels is asserted
    on the device (compute-sanitizer is not available on the GPU pool); a violation traps."""
    chk = os.path.join(ROOT, "tadbit_b200", "variants", "check.so")
    if not os.path.exists(chk):
        pytest.skip("variants/check.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    lines = _run({"TADBIT_B200_LIB": chk})
    assert all("'T': 0" not in l for l in lines)
    # the assertions are live: the library reports it
    import ctypes
    assert ctypes.CDLL(chk).tb_version() == 200 + 1000


@pytest.mark.gpu
def test_kernel_coverage_without_tiles():
    """TB_TILES=0 (A/B switch): sparse cells stay per-cell chunks, large cells of dense chunks
    stay in their overflow lists -- the list walk at the end of a dense group gets real work."""
    lines = _run({"TB_TILES": "0"})
    assert all("'T': 0" in l and "tile_cells=0" in l for l in lines)
    assert any("overflow=0" not in l for l in lines)
