"""Compile libb200dp.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).
One object per translation unit (compiled in parallel), linked into one shared library."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libb200dp.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--compiler-bindir", "/usr/bin/g++", "-Xcompiler", "-fPIC",
]


def units():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))
            if f.endswith((".cu", ".cuh", ".h"))] + \
        [os.path.join(os.path.dirname(HERE), "include", "b200dp.h")]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("B200DP_NVCC_EXTRA", "").split()      # e.g. -DB200DP_DEBUG_WAIT (diagnostic builds)
    os.makedirs(OBJ, exist_ok=True)
    newest_hdr = max(os.path.getmtime(s) for s in sources() if not s.endswith(".cu"))

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if not force and not extra and os.path.exists(obj) and \
                os.path.getmtime(obj) > max(newest_hdr, os.path.getmtime(src)):
            return obj, ""
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        r = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}")
        return obj, r.stdout

    with ThreadPoolExecutor(max_workers=4) as ex:
        results = list(ex.map(compile_one, units()))
    if verbose:
        for _, out in results:
            print(out)
    objs = [o for o, _ in results]
    subprocess.check_call([nvcc, "-shared", "-cudart", "static", "--compiler-bindir", "/usr/bin/g++",
                           "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
