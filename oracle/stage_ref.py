"""Stage the reference's own source files for the frame-encoding path into oracle/_ref/ (git-ignored, NOT
gpurun-ignored), so that the GPU box — where /root/reference does not exist — can time the UNMODIFIED
`models/vit.py` on its host cores (`bench.py --impl reference`, `cpu_baseline.kind = "reference"`).

    python -m oracle.stage_ref            # run in the build container; __graft_entry__.build() calls it

Nothing is copied into the git history: oracle/_ref/ is listed in .gitignore.  A manifest with the SHA-256 of every
staged file is written next to them, and oracle/reference_shims.py imports the files exactly as staged (the only
stand-ins are the timm / fairscale module shims, which carry no arithmetic except timm's PatchEmbed).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

SRC_ROOT = "/root/reference"
DST_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ["models/vit.py", "models/med.py"]  # the files oracle/reference_shims.py imports


def stage(verbose: bool = True) -> bool:
    if not os.path.isfile(os.path.join(SRC_ROOT, FILES[0])):
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC_ROOT, rel), os.path.join(DST_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST_ROOT, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC_ROOT, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"staged {len(FILES)} reference files into {DST_ROOT}")
    return True


if __name__ == "__main__":
    if not stage():
        raise SystemExit(f"{SRC_ROOT} not found: nothing staged")
