"""Import shim for the UNMODIFIED reference hot path (test infrastructure only).

Loads ``/root/reference/_pytadbit`` as package ``pytadbit`` without running its
``__init__`` (which needs compiled extensions and missing third-party modules).
It is used by ``oracle/gen_golden.py`` to produce the committed fixtures under
``tests/golden/``, by the optional ``tests/test_oracle_vs_reference.py`` (skipped when
the reference tree is absent) and by ``bench.py --impl reference`` to time the
unmodified reference on a small block.  In the build container the tree is
``/root/reference``; on the GPU box it is the untouched copy that
``__graft_entry__.build()`` places under ``baseline/_ref`` (git-ignored).

Recipe follows SURVEY.md Appendix A.
"""
import os
import sys
import types
import warnings

def _root():
    """/root/reference in the build container; on the GPU box the unmodified copy of the pure-Python
    package that __graft_entry__.build() leaves under baseline/_ref (git-ignored, it travels with
    the snapshot like the built libraries)."""
    env = os.environ.get("TADBIT_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/_pytadbit"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


REFERENCE_ROOT = _root()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "_pytadbit"))


def _stub(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Return a namespace with the reference's HiC_data, read_matrix and the four
    module-level hot-path functions."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    warnings.filterwarnings("ignore")
    sl = _stub("future.standard_library", install_aliases=lambda: None)
    _stub("future", standard_library=sl)
    _stub("pysam", AlignmentFile=object, __version__="0.22.0")
    pkg = _stub("pytadbit")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "_pytadbit")]
    from pytadbit.hic_data import HiC_data
    pkg.HiC_data = HiC_data
    _stub("pytadbit.parsers.hic_bam_parser", get_matrix=None)
    from pytadbit.parsers.hic_parser import read_matrix
    from pytadbit.utils.normalize_hic import iterative, expected
    from pytadbit.utils.hic_filtering import filter_by_mean, filter_by_zero_count
    ns = types.SimpleNamespace(HiC_data=HiC_data, read_matrix=read_matrix,
                               iterative=iterative, expected=expected,
                               filter_by_mean=filter_by_mean,
                               filter_by_zero_count=filter_by_zero_count,
                               root=REFERENCE_ROOT)
    _loaded["ns"] = ns
    return ns
