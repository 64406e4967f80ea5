"""Build the CUDA shared library in-tree with nvcc for sm_100a (no torch headers, plain C ABI).

Every .cu is compiled to its own object (in parallel) and the objects are linked into
`marl_sap_b200/libmarl_sap_b200.so`.  An object is reused only when the SHA-256 of its source, of every header under
`csrc/` and `include/`, and of the compiler flags matches the stamp written next to it, so a stale library can not
survive a source change (mtimes are not trusted: a fresh checkout gives every file the same one).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
# SAP_ABLATE=1 builds (and _lib.py then loads) a SEPARATE library with the profiling hooks, next to the release one
ABLATE = os.environ.get("SAP_ABLATE") == "1"
OBJ = os.path.join(PKG, "build_ablate" if ABLATE else "build")
LIB_PATH = os.path.join(PKG, "libmarl_sap_b200_ablate.so" if ABLATE else "libmarl_sap_b200.so")
STAMP = os.path.join(OBJ, "libmarl_sap_b200.stamp")
SOURCES = ["sap_real.cu", "sap_real_fast.cu", "sap_real_fast2.cu", "sap_real_large.cu", "sap_mock.cu", "sap_select.cu",
           "sap_lsa.cu", "sap_buffer.cu", "sap_power.cu", "sap_proximity.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _extra_flags() -> list[str]:
    # SAP_ABLATE=1 compiles the timing-ablation / forced-path hooks into the kernels (profiling builds only)
    return ["-DSAP_ABLATE=1"] if ABLATE else []


def _headers() -> list[str]:
    hs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    inc = os.path.join(ROOT, "include")
    return hs + [os.path.join(inc, f) for f in sorted(os.listdir(inc)) if f.endswith(".h")]


def _digest(paths: list[str], extra: str = "") -> str:
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as fh:
            h.update(p.encode() + b"\0" + fh.read())
    return h.hexdigest()


def _read(path: str) -> str:
    try:
        with open(path) as fh:
            return fh.read().strip()
    except OSError:
        return ""


def source_digest() -> str:
    """Digest of everything the library is built from."""
    return _digest([os.path.join(CSRC, s) for s in SOURCES] + _headers(), " ".join(FLAGS + _extra_flags()))


def needs_build() -> bool:
    return not os.path.exists(LIB_PATH) or _read(STAMP) != source_digest()


def _compile_one(src: str, verbose: bool) -> tuple[str, str]:
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, src[:-3] + ".o")
    want = _digest([path] + _headers(), " ".join(FLAGS + _extra_flags()))
    if os.path.exists(obj) and _read(obj + ".stamp") == want:
        return obj, ""
    cmd = [nvcc_path(), *FLAGS, *_extra_flags(), "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", path, "-o", obj]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    with open(obj + ".stamp", "w") as fh:
        fh.write(want)
    return obj, proc.stdout + proc.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            if f.endswith(".stamp"):
                os.remove(os.path.join(OBJ, f))
    with cf.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda s: _compile_one(s, verbose), SOURCES))
    cmd = [nvcc_path(), "--shared", "-o", LIB_PATH] + [o for o, _ in results]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    with open(STAMP, "w") as fh:
        fh.write(source_digest())
    if verbose:
        print("".join(log for _, log in results))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
