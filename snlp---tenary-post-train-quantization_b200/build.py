"""Build libtq100.so (sm_100a only) in-tree with nvcc.  No torch extension machinery: the library
is a plain C-ABI shared object (include/tq100.h) that the Python mirror loads with ctypes."""

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_build")
LIB = os.path.join(PKG, "libtq100.so")
SOURCES = ["api", "hessian_ffma", "hessian_tc", "cholinv", "atq", "ssr", "feedback", "codec", "sweep", "comm", "gemm_tc", "ternary_linear", "ternary_gemm_tc"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(PKG), "include", "tq100.h"))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()

    def compile_one(name):
        src = os.path.join(CSRC, name + ".cu")
        obj = os.path.join(OBJ, name + ".o")
        if not force and not _stale(obj, [src] + hdrs):
            return name, 0, ""
        cmd = [NVCC] + FLAGS + ["-Xptxas", "-v", "-c", src, "-o", obj]
        p = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(OBJ, name + ".log"), "w") as f:
            f.write(p.stdout + p.stderr)
        return name, p.returncode, p.stdout + p.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    failed = [(n, out) for n, rc, out in results if rc != 0]
    if failed:
        for n, out in failed:
            sys.stderr.write(f"--- nvcc failed for {n}.cu ---\n{out}\n")
        raise RuntimeError("libtq100 build failed: " + ", ".join(n for n, _ in failed))
    objs = [os.path.join(OBJ, n + ".o") for n in SOURCES]
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("libtq100 link failed:\n" + p.stdout + p.stderr)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
