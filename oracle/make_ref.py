"""Materialise the UNMODIFIED reference under oracle/_ref/ so that it can travel to the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY (never imported by the product package).  `/root/reference` exists only in the build
container; `oracle/_ref/` is git-ignored (the reference's sources never enter this repo's history) but NOT
gpurun-ignored, so the copy made here rides along with the snapshot.  What is copied: the Python package tree
`/root/reference/src` (sources and the yaml configs the runners read), byte for byte - no file is edited.  It is used by
  * `bench.py --impl reference` / the `cpu_baseline` leg: the reference's own ParallelRunner + BasicMAC + RNNAgent +
    epsilon-greedy selector on the box's host cores (imported through the stubs of `oracle/ref_import.py`);
  * `tests/test_gpu_dropin.py`: the reference's `run.run_sequential` driving this repo's runner / MAC / ReplayBuffer.

    python oracle/make_ref.py            # (re)build oracle/_ref from /root/reference
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC_ROOT = os.environ.get("MARL_SAP_REFERENCE", "/root/reference")
KEEP_EXT = (".py", ".yaml", ".yml", ".txt", ".md")


def tree_digest(root: str) -> str:
    h = hashlib.sha256()
    for d, _, files in sorted(os.walk(root)):
        for f in sorted(files):
            if f.endswith(KEEP_EXT):
                p = os.path.join(d, f)
                h.update(os.path.relpath(p, root).encode() + b"\0")
                with open(p, "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False) -> str | None:
    src = os.path.join(SRC_ROOT, "src")
    if not os.path.isdir(os.path.join(src, "envs")):
        return DEST if os.path.isdir(os.path.join(DEST, "src", "envs")) else None  # GPU box: use what travelled
    want = tree_digest(src)
    stamp = os.path.join(DEST, "SOURCE_SHA256")
    if not force and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return DEST
    shutil.rmtree(DEST, ignore_errors=True)
    for d, _, files in os.walk(src):
        if "__pycache__" in d:
            continue
        for f in files:
            if f.endswith(KEEP_EXT):
                rel = os.path.relpath(os.path.join(d, f), SRC_ROOT)
                out = os.path.join(DEST, rel)
                os.makedirs(os.path.dirname(out), exist_ok=True)
                shutil.copyfile(os.path.join(d, f), out)
    for f in ("LICENSE", "NOTICE", "requirements.txt"):
        if os.path.exists(os.path.join(SRC_ROOT, f)):
            shutil.copyfile(os.path.join(SRC_ROOT, f), os.path.join(DEST, f))
    with open(stamp, "w") as fh:
        fh.write(want + "\n")
    return DEST


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print(out if out else "reference tree not found (neither /root/reference nor oracle/_ref)")
