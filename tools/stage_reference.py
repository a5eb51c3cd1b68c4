#!/usr/bin/env python
"""Stage the UNMODIFIED reference Python the checker and the reference bench arm need under baseline/_ref/ (git-ignored,
travels to the GPU box with the snapshot; /root/reference does not exist there).

    python tools/stage_reference.py            # run in the build container; __graft_entry__.build() calls it too

Copies, byte for byte: lib/flownet/{model,networks,utils}/**/*.py and lib/pose/{models,utils}/**/*.py.  Nothing under
baseline/_ref is product code: only tests/, bench.py's reference arm / reference_gpu leg and oracle/ref_loader.py read it."""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
TREES = ["lib/flownet/model", "lib/flownet/networks", "lib/flownet/utils", "lib/pose/models", "lib/pose/utils"]
EXTRA = ["lib/flownet/__init__.py", "lib/pose/__init__.py"]


def stage():
    if not os.path.isdir(REF):
        return None
    manifest = {}
    for tree in TREES:
        for dirpath, _dirs, files in os.walk(os.path.join(REF, tree)):
            for f in files:
                if f.endswith(".py"):
                    src = os.path.join(dirpath, f)
                    rel = os.path.relpath(src, REF)
                    dst = os.path.join(DST, rel)
                    os.makedirs(os.path.dirname(dst), exist_ok=True)
                    shutil.copyfile(src, dst)
                    manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    for rel in EXTRA:
        src = os.path.join(REF, rel)
        if os.path.exists(src):
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1, sort_keys=True)
    return manifest


if __name__ == "__main__":
    m = stage()
    print("staged %d reference files under %s" % (len(m), DST) if m is not None else "no /root/reference here: nothing staged")
    sys.exit(0)
