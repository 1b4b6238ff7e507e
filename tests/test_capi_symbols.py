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


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/tadbit_b200.h is a C header (no C++ or torch types): a C99 program that includes it
    compiles with -pedantic, links against the library and calls entry points that need no GPU --
    what a cgo / JNI / N-API binding of the reference's maintainers would start from."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "tadbit_b200")
    src = tmp_path / "use_header.c"
    src.write_text(
        '#include <stdio.h>\n#include "tadbit_b200.h"\n'
        'int main(void) {\n'
        '  int64_t sums[3] = {900, 30, 500}, cells[3] = {3, 2, 1}, nkeys = 0;\n'
        '  double out[5];\n'
        '  if (tb_version() < 100) return 2;\n'
        '  if (tb_pool_expected(3, 3, sums, cells, 400.0, out, &nkeys) != TB_OK) return 3;\n'
        '  printf("%lld %.17g %.17g %.17g\\n", (long long)nkeys, out[0], out[1], out[2]);\n'
        '  if (tb_pool_expected(3, 3, sums, cells, 400.0, NULL, &nkeys) != TB_ERR_ARG) return 4;\n'
        '  return tb_last_error()[0] ? 0 : 5;\n}\n')
    exe = str(tmp_path / "use_header")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I",
                           os.path.join(root, "include"), str(src), "-o", exe, "-L", lib_dir,
                           "-l:libtadbit_b200.so", "-Wl,-rpath," + lib_dir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out
    # distance 0 alone (900 > 400): 300; distances 1 and 2 pooled (30 + 500): 530 / 3; trailing key
    assert out.stdout.split() == ["4", "300", repr(530 / 3.0), repr(530 / 3.0)]
