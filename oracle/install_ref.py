"""Install the UNMODIFIED reference files the CPU baseline needs into git-ignored `baseline/_ref/`.

    python -m oracle.install_ref            (also called by __graft_entry__.build())

Test / measurement infrastructure, not product code. The reference is pure Python without a build
system (no setup.py / pyproject.toml), so "installing" it is copying the three files of the hot path
byte for byte: criteria.py, metrics.py, network/Dorn.py. The
copy is NEVER committed (`baseline/_ref/` is in .gitignore) but travels to the GPU box with the
repository snapshot, where `/root/reference` does not exist; `oracle/_ref_loader.py` loads the modules
by file path from whichever of the two roots is present. No-op when `/root/reference` is absent.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("MDE_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ("criteria.py", "metrics.py", os.path.join("network", "Dorn.py"))


def install(verbose: bool = False) -> bool:
    """Returns True when baseline/_ref holds the files afterwards."""
    if not os.path.isfile(os.path.join(SRC, "criteria.py")):
        return all(os.path.isfile(os.path.join(DST, f)) for f in FILES)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    if verbose:
        print("reference files installed under %s" % DST)
    return True


if __name__ == "__main__":
    ok = install(verbose=True)
    raise SystemExit(0 if ok else 1)
