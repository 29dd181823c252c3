"""ORACLE BUILD — test / bench infrastructure only, never part of the product path.

    python oracle/build.py

Places an UNMODIFIED copy of the reference (/root/reference, present in the build container only)
under oracle/_ref/ so that it travels to the GPU box with the gpurun snapshot (oracle/_ref/ is
git-ignored: reference sources never enter this repository's history, and it is NOT listed in
.gpurunignore).  The reference is pure Python — there is nothing to compile; "building" it is a
byte-for-byte copy plus a manifest of SHA-256 digests that `verify()` re-checks on the box.

Users of oracle/_ref (and nothing else):
  * bench.py --impl reference / reference-cuda and the cpu_baseline leg: the reference's own
    Attacker.get_gradient_eps_mat + the .norm().item() loop (attacker.py:100-108, 220-229);
  * tests/test_gpu_reference_dropin.py: linkteller_b200.patch.install(reference_path=oracle/_ref)
    under the reference's real main.py → GCNTrainer.eval_output → Attacker chain.
When /root/reference is absent (the GPU box) an existing oracle/_ref is left as it is.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(DST, "MANIFEST.json")


def _files(root):
    out = []
    for dp, dn, fn in os.walk(root):
        dn[:] = sorted(d for d in dn if d not in (".git", "__pycache__"))
        for f in sorted(fn):
            if f.endswith((".pyc", ".pyo")) or f == "MANIFEST.json":
                continue
            out.append(os.path.relpath(os.path.join(dp, f), root))
    return out


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def build(verbose: bool = False) -> str | None:
    """Copy /root/reference → oracle/_ref (only when the source tree exists).  Returns the path or None."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    files = _files(SRC)
    want = {f: _sha(os.path.join(SRC, f)) for f in files}
    if os.path.exists(MANIFEST):
        try:
            if json.load(open(MANIFEST))["files"] == want and all(os.path.exists(os.path.join(DST, f)) for f in files):
                return DST
        except Exception:
            pass
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for f in files:
        os.makedirs(os.path.dirname(os.path.join(DST, f)) or DST, exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    with open(MANIFEST, "w") as fh:
        json.dump({"source": SRC, "files": want}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"oracle/_ref: {len(files)} files copied from {SRC}")
    return DST


def verify() -> bool:
    """True when oracle/_ref exists and every file still has the digest recorded at copy time."""
    if not os.path.exists(MANIFEST):
        return False
    want = json.load(open(MANIFEST))["files"]
    return all(os.path.exists(os.path.join(DST, f)) and _sha(os.path.join(DST, f)) == d for f, d in want.items())


if __name__ == "__main__":
    p = build(verbose=True)
    print(p if p else "oracle/_ref: /root/reference not present and no earlier copy — nothing to do")
    if p and not verify():
        sys.exit("oracle/_ref does not match its manifest")
