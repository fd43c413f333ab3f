"""Stages the UNMODIFIED reference packages the benchmark's reference arms import into the
git-ignored `baseline/_ref/` (it travels to the GPU box with the repo snapshot; `/root/reference`
does not exist there). Only whole files are copied, byte for byte, and only into `baseline/_ref/`;
nothing under it is tracked or edited. Run by `__graft_entry__.build()` in the build container.

    python baseline/stage_ref.py            # copies model/ optim/ utils/ data/ config/
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
PACKAGES = ("model", "optim", "utils", "data", "config")


def stage(ref_root=None, verbose=False):
    """Returns DEST when the staged copy exists (fresh or already present), else None."""
    ref_root = ref_root or os.environ.get("HERO_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref_root, "model")):
        return DEST if os.path.isdir(os.path.join(DEST, "model")) else None
    os.makedirs(DEST, exist_ok=True)
    for pkg in PACKAGES:
        src, dst = os.path.join(ref_root, pkg), os.path.join(DEST, pkg)
        if not os.path.isdir(src):
            continue
        os.makedirs(dst, exist_ok=True)
        for name in sorted(os.listdir(src)):
            s, d = os.path.join(src, name), os.path.join(dst, name)
            if not os.path.isfile(s) or not name.endswith((".py", ".json")):
                continue
            if not os.path.exists(d) or not filecmp.cmp(s, d, shallow=False):
                shutil.copyfile(s, d)
                if verbose:
                    print("staged", os.path.relpath(d, HERE))
    return DEST


if __name__ == "__main__":
    out = stage(sys.argv[1] if len(sys.argv) > 1 else None, verbose=True)
    print("reference staged at", out)
