#!/usr/bin/env python
"""Recipe for oracle/_ref: a byte-for-byte copy of the reference's hot-path modules, so that the UNMODIFIED reference can
be timed (and used as a checker) on the GPU box, where /root/reference does not exist.

    python oracle/make_ref.py            (run by __graft_entry__.build() whenever /root/reference is present)

oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored: it travels
with the snapshot like the built .so files.  Copied: ransac.py, homography.py (the two modules of the hot path), the
fixture matchespoints.npy and LICENSE (MIT).  A manifest with SHA-256 digests records what was copied.  Test
infrastructure only: nothing in ransac_with_homography_b200/ imports it."""
import hashlib
import json
import os
import shutil
import sys

REF = os.environ.get("REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["ransac.py", "homography.py", "matchespoints.npy", "LICENSE"]


def main():
    if not os.path.isdir(REF):
        print(f"{REF} not present: oracle/_ref left as it is")
        return 0
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(DST, f))
        with open(os.path.join(DST, f), "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": "choice17/ransac_with_homography (unmodified)", "sha256": manifest}, fh, indent=1)
    print("oracle/_ref:", ", ".join(FILES))
    return 0


if __name__ == "__main__":
    sys.exit(main())
