"""Build libhvb.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.

    python hockey-vision-analytics_b200/build.py [--force] [--verbose]

Output: hockey-vision-analytics_b200/hvb/libhvb.so (git-ignored; travels to the GPU box with gpurun).
Flags: no --use_fast_math and --fmad=false — IEEE division / exp and un-contracted mul+add are
needed for the bit-exact and identical-keep-set parity requirements.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "hvb", "libhvb.so")
INCLUDE = os.path.join(ROOT, "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--fmad=false",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
    "-Xptxas", "-v",
    "-I", INCLUDE, "-I", CSRC,
] + os.environ.get("HVB_NVCC_EXTRA", "").split()
if os.environ.get("HVB_BUILD_TAG"):            # A/B variants: separate objects and library name
    OBJ = OBJ + "_" + os.environ["HVB_BUILD_TAG"]
    OUT = os.path.join(HERE, "hvb", "libhvb_%s.so" % os.environ["HVB_BUILD_TAG"])


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libhvb cannot be built (there is no CPU fallback)")


def _newer(src_paths, target) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(p) > t for p in src_paths)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc"))]
    headers.append(os.path.join(INCLUDE, "hvb.h"))
    headers.append(os.path.abspath(__file__))

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        path = os.path.join(CSRC, src)
        if not force and not _newer([path] + headers, obj):
            return src, "", False
        cmd = [nvcc] + NVCC_FLAGS + ["-c", path, "-o", obj]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, p.stdout, p.stderr))
        with open(os.path.join(OBJ, src[:-3] + ".ptxas.log"), "w") as f:
            f.write(p.stderr)
        return src, p.stderr, True

    rebuilt = False
    with cf.ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        for src, log, did in ex.map(compile_one, sources):
            rebuilt |= did
            if verbose and did:
                print("== %s\n%s" % (src, log))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in sources]
    if force or rebuilt or _newer(objs, OUT):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
               "-Xcompiler", "-fPIC", "-o", OUT] + objs
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
