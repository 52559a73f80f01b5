"""Recipe for `oracle/_ref/`: the reference's own Python sources, staged VERBATIM — TEST INFRASTRUCTURE ONLY.

    python -m oracle.build_ref            (build container only: needs /root/reference, read-only)

The reference is pure Python over stock torch CPU operators, so "building" it is a file copy of `src/**/*.py`
from where it lies (`/root/reference/src`) into `oracle/_ref/src`.  `oracle/_ref/` is git-ignored (no reference
source enters the history) but NOT gpurun-ignored, so the staged copy travels to the GPU box exactly like the
built `.so` does, where `bench.py --impl reference` / `cpu_baseline` run it behind the `oracle/shims.py`
stand-ins for its uninstalled third-party imports (`kind: "reference"`).  When `oracle/_ref` is absent (a box
that never saw /root/reference) those legs fall back to the oracle port (`kind: "port"`).
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_SRC = os.path.join(REF_DIR, "src")
UPSTREAM = "/root/reference/src"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "entities", "algorithms", "ppo.py"))


def build_ref(force: bool = False) -> bool:
    """Stage the reference sources; returns whether `oracle/_ref` is usable afterwards."""
    if not os.path.isdir(UPSTREAM):
        return available()
    if available() and not force:
        return True
    if os.path.isdir(REF_DIR):
        shutil.rmtree(REF_DIR)
    n = 0
    for root, dirs, files in os.walk(UPSTREAM):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        rel = os.path.relpath(root, UPSTREAM)
        for f in files:
            if f.endswith(".py"):
                os.makedirs(os.path.join(REF_SRC, rel), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), os.path.join(REF_SRC, rel, f))
                n += 1
    with open(os.path.join(REF_DIR, "STAGED_FROM"), "w") as fh:
        fh.write(f"{UPSTREAM} ({n} files, verbatim)\n")
    return available()


if __name__ == "__main__":
    print("oracle/_ref staged" if build_ref(force=True) else "no /root/reference here: oracle/_ref not staged")
