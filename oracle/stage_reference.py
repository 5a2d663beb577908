"""ORACLE -- test infrastructure only.  Recipe that stages the UNMODIFIED reference next to the oracle.

    python oracle/stage_reference.py            (also run by __graft_entry__.build())

The reference is a pure-Python tree, so there is nothing to compile: the two files that hold the hot path
(SSD_from_scratch.py and SSD_trainer.py) are byte-copied from /root/reference -- which only exists in
the build container -- into oracle/_ref/, together with a manifest of their sha256.  oracle/_ref/ is git-ignored (the
sources never enter this repository's history) but not gpurun-ignored, so the staged files travel to the GPU box with the
built libssdhot.so.  There they serve as

  * the timed CPU arm (`bench.py --impl reference`, `cpu_baseline`: kind = "reference"), and
  * the device-matched checker of the `-m gpu` tests (the same functions called with device='cuda').

Nothing under ssdhot/ imports them (tests/test_host_cpu.py::test_product_never_imports_the_oracle).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
FILES = ("SSD_from_scratch.py", "SSD_trainer.py")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(verbose: bool = True) -> bool:
    """Copy the reference files if the reference tree is mounted; -> True if oracle/_ref is usable afterwards."""
    if all(os.path.isfile(os.path.join(REF_SRC, f)) for f in FILES):
        os.makedirs(REF_DST, exist_ok=True)
        manifest = {}
        for f in FILES:
            shutil.copyfile(os.path.join(REF_SRC, f), os.path.join(REF_DST, f))
            manifest[f] = _sha(os.path.join(REF_DST, f))
        with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as f:
            json.dump({"source": REF_SRC, "sha256": manifest}, f, indent=1)
        if verbose:
            print(f"[oracle] staged {', '.join(FILES)} -> {REF_DST}")
        return True
    ok = all(os.path.isfile(os.path.join(REF_DST, f)) for f in FILES)
    if verbose:
        print(f"[oracle] {REF_SRC} not mounted; " + ("using the staged copy" if ok else "no staged copy either"))
    return ok


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
