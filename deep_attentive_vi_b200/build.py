"""Build the C-ABI shared library (libdavi.so) in-tree with nvcc for sm_100a.

No torch types cross the boundary, so this is a plain ``nvcc -shared`` of
``csrc/*.cu`` against ``include/davi.h``; the result lands next to this file so
it travels with the repo snapshot to the GPU box.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdavi.so")
STAMP = os.path.join(HERE, "libdavi.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xcompiler", "-fPIC",
    "-shared",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [
        os.path.join(ROOT, "include", "davi.h")]


def _digest(paths, extra=""):
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:        # repo-relative names: the snapshot on the GPU box lives under another root
            h.update(os.path.relpath(p, ROOT).encode() + b"\0" + f.read())
    h.update((" ".join(NVCC_FLAGS) + extra).encode())
    return h.hexdigest()


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build(force=False, verbose=False):
    """Compile libdavi.so if sources changed: one object per .cu (compiled in parallel, each cached under
    build/ by the digest of the file + all headers), then one link.  Returns the library path."""
    srcs, hdrs = _sources(), _headers()
    dig = _digest(srcs + hdrs)
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == dig:
                return LIB
    nvcc = nvcc_path()
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libdavi.so")
    defines = []
    if os.path.exists(os.path.join(CSRC, "nonlocal_tc.cu")):
        defines.append("-DDAVI_HAVE_NONLOCAL_TC")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + defines + ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        cflags += ["-Xptxas", "-v"]

    def compile_one(src):
        d = _digest([src] + hdrs, " ".join(defines))
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + "." + d[:16] + ".o")
        if force or verbose or not os.path.exists(obj):
            cmd = [nvcc] + cflags + ["-c", src, "-o", obj + ".tmp"]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose:
                sys.stderr.write(res.stderr)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
            os.replace(obj + ".tmp", obj)
        return obj

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, srcs))
    keep = set(objs)
    for f in os.listdir(objdir):                       # drop objects of older source versions
        if f.endswith(".o") and os.path.join(objdir, f) not in keep:
            os.remove(os.path.join(objdir, f))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
           "-o", LIB + ".tmp"] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    os.replace(LIB + ".tmp", LIB)
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
