"""TEST INFRASTRUCTURE ONLY -- stage the UNMODIFIED reference for the GPU box.

The reference (``/root/reference/pta_replicator``, six pure-Python modules) has nothing to compile with gcc; its
"binary" is CPython bytecode.  This recipe byte-compiles the modules FROM WHERE THEY LIE into ``oracle/_ref/``
(git-ignored, not gpurun-ignored: like a built ``.so`` it travels to the GPU box but stays out of the history; no
reference source is copied).  ``oracle/refstubs.py`` imports the package from ``/root/reference`` when that exists
and from ``oracle/_ref`` (sourceless ``.pyc`` modules, same interpreter image) otherwise, so ``bench.py --impl
reference`` and the ``cpu_baseline`` leg time the reference's own functions on the box's host cores.

    python -m oracle.build_ref        # also run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import hashlib
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/pta_replicator"
OUT = os.path.join(HERE, "_ref")
MODULES = ("__init__", "constants", "white_noise", "red_noise", "deterministic", "spharmORFbasis", "simulate")


def build(force: bool = False) -> bool:
    """Byte-compile the reference package into ``oracle/_ref/pta_replicator``.  Returns True when staged."""
    if not os.path.isdir(REF_SRC):
        return os.path.isfile(os.path.join(OUT, "MANIFEST.json"))
    pkg = os.path.join(OUT, "pta_replicator")
    os.makedirs(pkg, exist_ok=True)
    manifest = {"python": sys.version.split()[0], "magic": py_compile.importlib.util.MAGIC_NUMBER.hex(), "modules": {}}
    for m in MODULES:
        src = os.path.join(REF_SRC, m + ".py")
        dst = os.path.join(pkg, m + ".pyc")
        with open(src, "rb") as fh:
            digest = hashlib.sha256(fh.read()).hexdigest()
        manifest["modules"][m] = digest
        if force or not os.path.isfile(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            # unchecked-hash pyc: valid without the source file next to it
            py_compile.compile(src, cfile=dst, dfile=f"<reference>/pta_replicator/{m}.py", doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    return True


if __name__ == "__main__":
    print("staged" if build(force="--force" in sys.argv) else "reference not available", OUT)
