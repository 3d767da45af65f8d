#!/usr/bin/env python
"""Stage the UNMODIFIED reference backbone under baseline/_ref/ (git-ignored, travels to the GPU box with gpurun) so that
`bench.py --impl reference` can time the reference's own module there.

The reference (alexanderswerdlow/unidisc) is not a pip-installable package (no setup.py / pyproject build target for the
trainer; its pinned torch 2.6+cu124 cannot run on sm_100), so the "install" is a verbatim copy of the files the stock
`models.dit.DIT` imports — nothing is edited:

  models/__init__.py  models/dit.py  models/standalone_rotary.py  models/noise_schedule.py
  decoupled_utils.py  unidisc/utils/tensor_utils.py

The two absent third-party imports (omegaconf, diffusers' Lumina 2-D RoPE helper) are shimmed at import time by
oracle/ref_loader.py, exactly as for the golden fixtures.  Run in the build container:
    python baseline/install_reference.py            (also called by __graft_entry__.build() when /root/reference exists)
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["models/__init__.py", "models/dit.py", "models/standalone_rotary.py", "models/noise_schedule.py", "decoupled_utils.py",
         "unidisc/utils/tensor_utils.py"]


def install(src_root: str = "/root/reference") -> str:
    if not os.path.isfile(os.path.join(src_root, "models", "dit.py")):
        raise RuntimeError(f"reference not found at {src_root}")
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(src_root, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
            manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
        elif rel.endswith("__init__.py"):
            open(dst, "w").close()
    json.dump(dict(source=src_root, sha256=manifest), open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    return DST


if __name__ == "__main__":
    print(install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
