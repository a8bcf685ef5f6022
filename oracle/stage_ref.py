"""Stage the reference's own Python sources under oracle/_ref/ so that the reference arm of bench.py
(`--impl reference`) drives the REAL modeling_csm.py on the GPU box's host cores.

TEST / MEASUREMENT INFRASTRUCTURE.  oracle/_ref/ is git-ignored (reference sources never enter the history) but not
gpurun-ignored, so the staged copy travels with the snapshot.  Run in the build container, where /root/reference
exists; __graft_entry__.build() calls it.  Nothing in the product imports from oracle/.
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("CSM_REFERENCE_DIR", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = ("modeling_csm.py", "processor.py", "LICENSE")


def stage(verbose: bool = False) -> bool:
    """Copy the reference's files (unchanged) into oracle/_ref/.  Returns True when a staged copy is available."""
    if os.path.isfile(os.path.join(REF_SRC, "modeling_csm.py")):
        os.makedirs(REF_DST, exist_ok=True)
        for f in FILES:
            src = os.path.join(REF_SRC, f)
            if os.path.isfile(src):
                shutil.copyfile(src, os.path.join(REF_DST, f))
        if verbose:
            print(f"staged {REF_SRC} -> {REF_DST}")
    return os.path.isfile(os.path.join(REF_DST, "modeling_csm.py"))


if __name__ == "__main__":
    print("available" if stage(verbose=True) else "reference not found")
