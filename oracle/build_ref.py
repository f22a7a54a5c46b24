"""TEST INFRASTRUCTURE -- recipe that makes the UNMODIFIED reference module travel to the GPU box.

The reference is pure Python (no native code), so "building" it is a plain install of the two files
of the hot path from where they lie under ``/root/reference`` into ``oracle/_ref/f_lite/`` (git-ignored, NOT
gpurun-ignored: it ships with the snapshot like the built ``.so``, and never enters the history).  Nothing is edited;
``ORIGIN.json`` records the source path and the sha256 of every file so a reader can verify that.

Used by: ``oracle/ref_shim.py`` (falls back to this copy when ``/root/reference`` is absent), the ``-m gpu`` tests
that run the real module with real Liger / flash-attn kernels, and ``bench.py --impl reference``.
Run by ``__graft_entry__.build()`` in the build container; on the GPU box the prebuilt copy is used as is.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

REF_ROOT = "/root/reference"
FILES = ("f_lite/model.py", "f_lite/pipeline.py")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def ref_path(rel: str):
    """Path of a reference file: the mounted reference when present (build container), else the travelled copy."""
    for base in (REF_ROOT, OUT):
        p = os.path.join(base, rel)
        if os.path.exists(p):
            return p
    return None


def build(verbose: bool = True) -> bool:
    """Copy FILES into oracle/_ref/ (only when the mounted reference exists).  Returns True when a copy is in place."""
    if not os.path.isdir(REF_ROOT):
        return all(os.path.exists(os.path.join(OUT, f)) for f in FILES)
    origin = {}
    for rel in FILES:
        src, dst = os.path.join(REF_ROOT, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        data = open(src, "rb").read()
        if not os.path.exists(dst) or open(dst, "rb").read() != data:
            shutil.copyfile(src, dst)
        origin[rel] = {"source": src, "sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data)}
    with open(os.path.join(OUT, "ORIGIN.json"), "w") as f:
        json.dump(origin, f, indent=1)
    if verbose:
        print("[oracle/_ref]", ", ".join(f"{k} ({v['sha256'][:12]})" for k, v in origin.items()))
    return True


if __name__ == "__main__":
    build()
