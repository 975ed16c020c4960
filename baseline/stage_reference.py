#!/usr/bin/env python
"""Stage the UNMODIFIED reference (MingSun-Tse/Efficient-NeRF) under baseline/_ref/ so that it travels to the GPU box.

    python baseline/stage_reference.py [--src /root/reference] [--check]

`/root/reference` exists only in the build container; `gpurun` ships `/root/repo` (minus `.gpurunignore`), and
`baseline/_ref/` is git-ignored but NOT gpurun-ignored.  The reference is plain Python with no setup.py / pyproject,
so "installing" it is copying its sources byte for byte: every *.py, the configs and the licence — nothing is edited,
nothing enters the git history.  `baseline/MANIFEST.sha256` (tracked: hashes only) pins what was staged; `--check`
verifies a staged tree against it, which is how the GPU-side tests and `bench.py --impl reference` prove that the code
they run is the reference's own.

What uses the staged tree (never the product path):
  * `bench.py --impl reference`  — times the reference's model/nerf_raybased.py, utils/run_nerf_raybased_helpers.py and
    main.py's render functions on the box's host cores (`cpu_baseline.kind = "reference"`);
  * `tests/test_reference_main.py` — runs the unmodified main.py --render_only / --benchmark through
    `efficient_nerf_b200.dropin` on a B200.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(HERE, "MANIFEST.sha256")
KEEP_EXT = (".py", ".txt", ".md")
KEEP_NAMES = ("LICENSE",)


def _files(src):
    out = []
    for d, dirs, files in os.walk(src):
        dirs[:] = sorted(x for x in dirs if x not in (".git", "figs", "__pycache__"))
        for f in sorted(files):
            if f.endswith(KEEP_EXT) or f in KEEP_NAMES:
                out.append(os.path.relpath(os.path.join(d, f), src))
    return out


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(src="/root/reference", dst=DST, write_manifest=True):
    if not os.path.isdir(src):
        raise FileNotFoundError(f"reference checkout not found at {src}")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    man = {}
    for rel in _files(src):
        os.makedirs(os.path.dirname(os.path.join(dst, rel)) or dst, exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), os.path.join(dst, rel))
        man[rel] = _sha(os.path.join(dst, rel))
    if write_manifest:
        with open(MANIFEST, "w") as f:
            json.dump(man, f, indent=0, sort_keys=True)
            f.write("\n")
    return man


def check(dst=DST):
    """Raises unless every file of the manifest is present in `dst` with the recorded hash (and no other .py is)."""
    with open(MANIFEST) as f:
        man = json.load(f)
    bad = [rel for rel, h in man.items() if not os.path.exists(os.path.join(dst, rel)) or _sha(os.path.join(dst, rel)) != h]
    extra = [rel for rel in _files(dst) if rel not in man]
    if bad or extra:
        raise RuntimeError(f"baseline/_ref differs from the staged reference: changed/missing {bad[:5]}, extra {extra[:5]}")
    return len(man)


def staged_root(required=True):
    """baseline/_ref if it holds a verified copy, else $R2L_REFERENCE or /root/reference (build container)."""
    if os.path.exists(os.path.join(DST, "main.py")):
        check(DST)
        return DST
    for cand in (os.environ.get("R2L_REFERENCE"), "/root/reference"):
        if cand and os.path.exists(os.path.join(cand, "main.py")):
            return cand
    if required:
        raise FileNotFoundError("no reference tree: run `python baseline/stage_reference.py` in the build container")
    return None


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=os.environ.get("R2L_REFERENCE", "/root/reference"))
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    if a.check:
        print(f"baseline/_ref: {check()} files match baseline/MANIFEST.sha256")
    else:
        m = stage(a.src)
        print(f"staged {len(m)} files from {a.src} into {DST}")
    sys.exit(0)
