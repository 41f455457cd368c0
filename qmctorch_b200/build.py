"""Build libqmcb.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m qmctorch_b200.build [--force]

Plain ``nvcc -shared``: no torch headers are involved (the ABI is raw pointers), so the
build takes about a minute and cross-compiles without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libqmcb.so")
SOURCES = ["plan.cu", "fused_psi.cu", "fused_eloc.cu", "fused_grad.cu", "fused_mh.cu", "operators.cu", "backward.cu"]
HEADERS = ["plan.h", "device.cuh", "philox.cuh", "fused_impl.cuh", os.path.join("..", "..", "include", "qmcb.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math=false",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force=False, verbose=False, extra_flags=(), out=None):
    """extra_flags / out: tuning experiments (e.g. -DQMCB_MINBLOCKS=3 into another file,
    selected at run time with QMCB_LIB=path)."""
    if out is None and not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + list(extra_flags)
    objs = []
    tag = "" if out is None else "." + os.path.basename(out)

    def compile_one(src):
        obj = os.path.join(LIBDIR, src.replace(".cu", tag + ".o"))
        cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + \
            ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    target = LIB if out is None else out
    cmd = [nvcc, "-shared", "-o", target] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if out is not None:
        for o in objs:
            os.remove(o)
    return target


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    out = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")), None)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, extra_flags=extra, out=out))
