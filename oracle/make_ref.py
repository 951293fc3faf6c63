#!/usr/bin/env python
"""oracle/make_ref.py -- TEST / BENCH INFRASTRUCTURE ONLY.

Puts the UNMODIFIED reference selector path where the GPU box can run it: /root/reference does not exist there,
but ``oracle/_ref/`` (git-ignored, NOT gpurun-ignored) travels with the tree like the built ``.so`` files do.
The reference's selector modules are plain Python, so "building" them is copying: this script
imports the selectors from /root/reference exactly as ``reference_loader`` does, lists every ``det3d`` source
file that import pulled in (the selector modules, ``det3d/utils/registry.py``, ``det3d/torchie`` ...) and copies
those files, byte for byte and with their relative paths, into ``oracle/_ref/``.  No reference source enters the
repository's history.  ``__graft_entry__.build()`` calls ``main()`` whenever /root/reference is present.

``reference_loader`` falls back to ``oracle/_ref`` when /root/reference is absent, which is what makes
``bench.py --impl reference`` time the real reference (``cpu_baseline.kind == "reference"``) on the GPU box.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("ADSEL_REFERENCE_SRC", "/root/reference")

_LIST = r"""
import json, os, sys
sys.path.insert(0, {here!r})
sys.dont_write_bytecode = True
os.environ["ADSEL_REFERENCE_ROOT"] = {src!r}
import reference_loader as rl
rl.load_reference_selectors()
root = os.path.realpath({src!r}) + os.sep
files = sorted({{os.path.realpath(m.__file__) for m in list(sys.modules.values())
                if getattr(m, "__file__", None) and os.path.realpath(m.__file__).startswith(root)}})
print(json.dumps([f[len(root):] for f in files]))
"""


def main(verbose: bool = True) -> str:
    if not os.path.isfile(os.path.join(SRC, "det3d", "selectors", "spatial_temporal_selector.py")):
        raise RuntimeError(f"reference tree not found under {SRC}")
    out = subprocess.check_output([sys.executable, "-c", _LIST.format(here=HERE, src=SRC)], text=True,
                                  env={**os.environ, "PYTHONDONTWRITEBYTECODE": "1"})
    files = json.loads(out.strip().splitlines()[-1])
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    manifest = {}
    for rel in files:
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"oracle/_ref: {len(files)} unmodified reference files copied from {SRC}")
    return DEST


if __name__ == "__main__":
    main()
