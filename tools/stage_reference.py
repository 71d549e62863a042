#!/usr/bin/env python
"""Stage the UNMODIFIED reference where the GPU box can execute it.

``/root/reference`` does not exist on the GPU box; ``baseline/_ref/`` is
git-ignored (never enters the history) but not gpurun-ignored, so it travels
with the snapshot.  This copies ``dmc/`` byte for byte into ``baseline/_ref/dmc``
and writes ``baseline/_ref/MANIFEST.json`` (sha256 per file) so a reader can
check that what ran on the box is the stock code.  The drop-in tests
(``tests/test_gpu_dropin.py``) and ``tools/dropin_bench.py`` /
``tools/train_ddp_bench.py`` import it through ``oracle/load_reference.py``.

    python tools/stage_reference.py [--src /root/reference]
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=os.environ.get("DVC_REFERENCE_SRC", "/root/reference"))
    args = ap.parse_args()
    src = os.path.join(args.src, "dmc")
    if not os.path.isdir(src):
        print(f"stage_reference: {src} not found; nothing staged", file=sys.stderr)
        return 1
    dst_root = os.path.join(ROOT, "baseline", "_ref")
    dst = os.path.join(dst_root, "dmc")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst_root, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    manifest = {}
    for base, _, files in os.walk(dst):
        for f in sorted(files):
            p = os.path.join(base, f)
            manifest[os.path.relpath(p, dst_root)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    with open(os.path.join(dst_root, "MANIFEST.json"), "w") as fh:
        json.dump({"source": args.src, "files": manifest}, fh, indent=1, sort_keys=True)
    print(f"staged {len(manifest)} files from {src} -> {dst}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
