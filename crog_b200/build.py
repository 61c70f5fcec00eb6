"""In-tree build of libcrog_b200.so (hand-written sm_100a kernels behind the C ABI of
include/crog_b200.h).  nvcc cross-compiles without a GPU; the built .so is git-ignored but
travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(OUT_DIR, "obj")
SO_PATH = os.path.join(OUT_DIR, "libcrog_b200.so")
SOURCES = ["api.cu", "gemm_simt.cu", "gemm_tc.cu", "elementwise.cu", "attention.cu", "attention_tc.cu", "tail.cu", "ssg.cu", "warp.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
FLAGS.remove("--use_fast_math=false")
# same-box A/B of two builds: CROG_BUILD_EXTRA="-DNAME=value ..." adds compiler flags, CROG_BUILD_TAG=name builds into
# lib/libcrog_b200.<name>.so (objects in lib/obj.<name>/); load it with CROG_B200_SO=...
_TAG = os.environ.get("CROG_BUILD_TAG", "")
if os.environ.get("CROG_BUILD_EXTRA"):
    FLAGS += os.environ["CROG_BUILD_EXTRA"].split()
if _TAG:
    OBJ_DIR = os.path.join(OUT_DIR, "obj." + _TAG)
    SO_PATH = os.path.join(OUT_DIR, f"libcrog_b200.{_TAG}.so")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(PKG), "include", "crog_b200.h"))

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            r = subprocess.run([NVCC] + FLAGS + ["-c", s, "-o", o], capture_output=True, text=True)
            log = r.stdout + r.stderr
            with open(o + ".log", "w") as f:
                f.write(log)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{log}")
            if verbose:
                print(log, file=sys.stderr)
        return o

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(SO_PATH, objs):
        r = subprocess.run([NVCC, "-shared", "-o", SO_PATH] + objs + ["-cudart", "static"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return SO_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
