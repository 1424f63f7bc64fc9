"""Place the UNMODIFIED reference model file under the git-ignored ``oracle/_ref/`` (authoring container only).

    python oracle/build_ref.py

TEST INFRASTRUCTURE.  The reference's hot path is one pure-Python file (net/MP_HSIR.py, 856 lines; its only
non-installable imports, ``timm.models.layers`` and ``clip``, are stubbed by ``oracle/ref_import.py``).  It has no
build step, so "building" the real reference = a byte-for-byte copy of that file to ``oracle/_ref/net/MP_HSIR.py``
plus a SHA-256 manifest.  ``oracle/_ref/`` is listed in .gitignore (reference sources never enter the history) but
not in .gpurunignore, so the copy travels to the GPU box, where ``bench.py --impl reference`` times it on the host
cores at the full BASELINE shapes (``cpu_baseline.kind = "reference"``) and ``tests/test_reference_live.py``
cross-checks the committed goldens against it.  When /root/reference is absent (the GPU box) this script only
verifies an existing copy.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = "/root/reference"
DST_ROOT = os.path.join(HERE, "_ref")
FILES = ["net/MP_HSIR.py"]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(verbose: bool = False) -> bool:
    """-> True when oracle/_ref holds a verified copy."""
    man_path = os.path.join(DST_ROOT, "MANIFEST.json")
    if os.path.isdir(SRC_ROOT):
        man = {}
        for rel in FILES:
            src, dst = os.path.join(SRC_ROOT, rel), os.path.join(DST_ROOT, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not os.path.exists(dst) or _sha(dst) != _sha(src):
                shutil.copyfile(src, dst)
            man[rel] = _sha(src)
        with open(man_path, "w") as f:
            json.dump({"source": SRC_ROOT, "sha256": man}, f, indent=1)
    if not os.path.exists(man_path):
        return False
    with open(man_path) as f:
        man = json.load(f)["sha256"]
    ok = all(os.path.exists(os.path.join(DST_ROOT, rel)) and _sha(os.path.join(DST_ROOT, rel)) == h for rel, h in man.items())
    if verbose:
        print("oracle/_ref:", "verified" if ok else "CORRUPT", man)
    return ok


if __name__ == "__main__":
    sys.exit(0 if build(verbose=True) else 1)
