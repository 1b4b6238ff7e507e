"""Build the CUDA C-ABI library in-tree: tadbit_b200/libtadbit_b200.so (sm_100a only).

    python -m tadbit_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with gpurun.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtadbit_b200.so")
SOURCES = ["api.cu", "build.cu", "encode.cu", "rowsum.cu", "tiles.cu", "kernels.cu", "extras.cu", "compartments.cu", "synth.cu", "comm.cu", "textio.cu", "pooling.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--expt-relaxed-constexpr"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "tadbit_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """Compile the library.  ``defines``/``out`` build a tuning variant (e.g.
    ``defines=["TB_DENSE_WARPS=20"], out="variants/w20.so"``, loaded with TADBIT_B200_LIB)."""
    variant = out is not None
    target = os.path.join(HERE, out) if variant else LIB
    if not variant and not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    objdir = os.path.join(HERE, "build", os.path.basename(out) + ".d") if variant else os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    os.makedirs(os.path.dirname(target), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
        if verbose:
            sys.stderr.write(out)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", target] + objs + ["-ldl"]
    subprocess.check_call(cmd)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
