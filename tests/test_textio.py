"""tb_write_abc_lines (host-only): the cell lines of an .abc file, floats as Python's repr writes
them -- compared with Python itself on doubles of every magnitude, the special values and the
boundaries between fixed and exponent notation."""
import ctypes as C
import struct

import numpy as np

from tadbit_b200 import _capi


def _write(path, a, b, ints=None, vals=None):
    a = np.ascontiguousarray(a, dtype=np.int32)
    b = np.ascontiguousarray(b, dtype=np.int32)
    ip = vp = None
    if ints is not None:
        ints = np.ascontiguousarray(ints, dtype=np.int32)
        ip = _capi.ptr(ints, C.c_int32)
    if vals is not None:
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        vp = _capi.ptr(vals, C.c_double)
    _capi.check(_capi.lib().tb_write_abc_lines(str(path).encode(), a.size, _capi.ptr(a, C.c_int32),
                                               _capi.ptr(b, C.c_int32), ip, vp))


def test_float_text_is_pythons_repr(tmp_path):
    rng = np.random.default_rng(12)
    special = [0.0, -0.0, 1.0, -1.0, 0.1, 0.5, 1e-4, 9.999999999999999e-05, 1e-5, 0.00011, 1e15, 1e16,
               9999999999999998.0, 1e17, 123456789012345680.0, 1.5e300, 5e-324, 2.2250738585072014e-308,
               1.7976931348623157e308, float("inf"), float("-inf"), float("nan"), 1 / 3.0, 2 / 3.0, 100.0,
               1e22, 1e23, 4.35, 0.3, 2.675, 28.81836451240, 5e-5, 12345678.9, 0.001, 1234567890123456.7]
    bits = rng.integers(0, 2 ** 63, 300000, dtype=np.int64).astype(np.uint64) | \
        (rng.integers(0, 2, 300000).astype(np.uint64) << np.uint64(63))
    anyd = bits.view(np.float64)                             # every exponent, both signs, NaNs
    ratio = rng.integers(1, 5000, 300000) / rng.lognormal(0, 1, 300000) / rng.lognormal(0, 1, 300000)
    scaled = rng.random(200000) * 10.0 ** rng.integers(-8, 20, 200000)
    vals = np.concatenate([np.array(special), anyd, ratio, scaled])
    a = np.arange(vals.size) % 70000
    path = tmp_path / "f.abc"
    path.write_text("# header\n")
    _write(path, a, a[::-1], vals=vals)
    lines = path.read_text().splitlines()
    assert lines[0] == "# header" and len(lines) == vals.size + 1
    bad = []
    for k, (line, x) in enumerate(zip(lines[1:], vals.tolist())):
        want = "%d\t%d\t\t%r" % (a[k], a[::-1][k], x)
        if line != want:
            bad.append((line, want))
    assert not bad, bad[:5]
    # and they read back as the same doubles
    back = np.array([float(line.split("\t")[-1]) for line in lines[1:]])
    same = (back == vals) | (np.isnan(back) & np.isnan(vals))
    assert same.all()
    assert struct.pack("d", back[1]) == struct.pack("d", -0.0)


def test_integer_lines_append_and_errors(tmp_path):
    path = tmp_path / "r.abc"
    _write(path, [0, 5, 2147483647], [3, 5, 0], ints=[7, -2, 2147483647])
    _write(path, [], [], ints=[])
    _write(path, [1], [2], ints=[3])
    assert path.read_text() == "0\t3\t\t7\n5\t5\t\t-2\n2147483647\t0\t\t2147483647\n1\t2\t\t3\n"
    big = np.arange(100000, dtype=np.int32)                  # several write batches
    _write(tmp_path / "big.abc", big, big, ints=big)
    got = (tmp_path / "big.abc").read_text().splitlines()
    assert len(got) == big.size and got[-1] == "99999\t99999\t\t99999"
    try:
        _write(tmp_path / "no" / "such" / "dir.abc", [1], [1], ints=[1])
    except Exception as e:
        assert "cannot open" in str(e)
    else:
        raise AssertionError("writing into a missing directory must fail")
