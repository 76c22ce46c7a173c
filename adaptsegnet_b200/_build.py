"""Builds libasn_b200.so (sm_100a) in-tree with nvcc.  No GPU is needed to build."""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libasn_b200.so")
OBJ_DIR = os.path.join(ROOT, "build", "obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    return cand if os.path.exists(cand) else "nvcc"


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def source_id() -> str:
    """sha256 over every source the library is built from (csrc/*.cu, csrc/*.cuh, include/asn_b200.h, in name order) and
    the compiler flags: the identity of the committed sources.  It is compiled into the library (asn_build_id) so that
    the binary that ran can be matched against the tree (smoke() prints and checks it)."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))) + \
        [os.path.join(ROOT, "include", "asn_b200.h")]
    for f in files:
        h.update(os.path.basename(f).encode() + b"\0")
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(ROOT, "include", "asn_b200.h")]
    nvcc = _nvcc()
    sid = source_id()
    id_file = os.path.join(OBJ_DIR, "build_id.txt")
    last_id = open(id_file).read().strip() if os.path.exists(id_file) else ""
    # ASN_BUILD_FORCE=1 (or force=True): recompile everything from the sources in the tree, whatever the timestamps say
    force = force or os.environ.get("ASN_BUILD_FORCE") == "1"

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        is_capi = os.path.basename(src) == "capi.cu"      # carries the build id: rebuilt whenever any source changed
        if force or _stale(obj, [src] + headers) or (is_capi and sid != last_id):
            cmd = [nvcc] + NVCC_FLAGS + (["-DASN_BUILD_ID=\"%s\"" % sid] if is_capi else []) + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    if force or _stale(LIB_PATH, objs) or sid != last_id:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    with open(id_file, "w") as f:
        f.write(sid)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
