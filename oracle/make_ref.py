"""Stage the UNMODIFIED reference hot-path files for the GPU box -- TEST / BENCH INFRASTRUCTURE ONLY.

    python oracle/make_ref.py          (also run by __graft_entry__.build() when /root/reference is present)

The reference is pure Python, so "building" it means staging the two files the hot path lives in,
``src/periodicity/spectral.py`` and ``src/periodicity/phase.py``, byte for byte under ``oracle/_ref/periodicity/``
together with a manifest of their SHA-256 digests.  ``oracle/_ref/`` is git-ignored (it never enters the history:
reference sources are not part of this repository) but NOT gpurun-ignored, so it travels to the GPU box like the
built ``.so`` files do; there ``oracle/refload.py`` loads the files behind the numpy-only stand-in for
``periodicity.core`` and ``bench.py --impl reference`` / ``cpu_baseline`` time the reference's own code
(``kind: "reference"``).  Without the directory both fall back to the numpy port in ``oracle/`` (``kind: "port"``).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("PERIODICITY_REFERENCE_ROOT", "/root/reference")
FILES = ("spectral.py", "phase.py")


def main():
    src_dir = os.path.join(REFERENCE_ROOT, "src", "periodicity")
    if not all(os.path.isfile(os.path.join(src_dir, f)) for f in FILES):
        print(f"[make_ref] {src_dir} not found: nothing staged (oracle/_ref keeps whatever it holds)")
        return 0
    dst_dir = os.path.join(HERE, "_ref", "periodicity")
    os.makedirs(dst_dir, exist_ok=True)
    manifest = {"source": src_dir, "files": {}}
    for f in FILES:
        shutil.copyfile(os.path.join(src_dir, f), os.path.join(dst_dir, f))
        with open(os.path.join(dst_dir, f), "rb") as fh:
            manifest["files"][f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    print(f"[make_ref] staged {', '.join(FILES)} under {dst_dir}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
