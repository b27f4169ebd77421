"""Stages the two reference training scripts as TEST INPUT for tests/test_gpu_reference_scripts.py.

BASELINE.json north_star: "train_rna2dna.py and train_dna2rna.py run unmodified".  The reference tree exists only in the
build container (/root/reference); the GPU box gets a snapshot of this repository.  This script copies the two script files,
byte for byte, into tests/_ref_scripts/ -- a directory that is git-ignored (no reference source enters the history) but
travels with the snapshot, like oracle/_ref/ would for a compiled reference.  __graft_entry__.build() calls it whenever
/root/reference is present.  The test executes the copies by path with THIS repository's `src` package first on PYTHONPATH.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
DST = os.path.join(ROOT, "tests", "_ref_scripts")
SCRIPTS = ("train_rna2dna.py", "train_dna2rna.py")


def stage():
    if not os.path.isdir(REF):
        return False
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for name in SCRIPTS:
        src = os.path.join(REF, name)
        shutil.copyfile(src, os.path.join(DST, name))
        with open(src, "rb") as f:
            manifest[name] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "sha256": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    print("staged" if stage() else "no /root/reference here: nothing staged")
    sys.exit(0)
