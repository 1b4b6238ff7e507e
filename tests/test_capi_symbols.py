"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/tadbit_b200.h declares; compute calls fail loudly without a device."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "tadbit_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from tadbit_b200 import _capi
    lib = _capi.lib()
    names = _declared()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), name
        assert name in _capi.PROTOTYPES, "no ctypes prototype for %s" % name
    assert sorted(_capi.PROTOTYPES) == names
    assert lib.tb_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device the product path raises; it never computes on the host."""
    import tadbit_b200
    if tadbit_b200.device_count() > 0:
        pytest.skip("CUDA device present")
    with pytest.raises(tadbit_b200.B200Error, match="no CUDA device"):
        tadbit_b200.HiC_data.from_coo(4, [0, 1], [1, 2], [3, 4])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tadbit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", "").lower() or f == "__init__.py" \
                    or "import tadbit_oracle" not in src, f
                assert "tadbit_oracle" not in src, f
                assert "ref_shim" not in src, f
