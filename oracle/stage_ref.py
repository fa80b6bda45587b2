"""Stage the UNMODIFIED reference into the git-ignored oracle/_ref/ (test / baseline infrastructure).

    python oracle/stage_ref.py            (also run by __graft_entry__.build() where /root/reference exists)

The reference (russellgeum/Digging-into-Self-Supervised-Monocular-Depth-Estimation) is pure Python: there is
nothing to compile.  Its hot path (model_layer/warp.py, model_loss/model_loss.py, model_tool/processor.py:139-218)
imports the rest of its package tree at module top, so the recipe copies the Python packages as they lie under
/root/reference - byte for byte, nothing edited, nothing added - into oracle/_ref/, which is listed in .gitignore
(never part of the history) but not in .gpurunignore, so that it travels to the GPU box like a built .so.
`bench.py --impl reference` and the `gpu_reference` leg then time the reference's own code (oracle/ref_loader.py);
where oracle/_ref/ is absent they fall back to the port (oracle/oracle_torch.py) and say so.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
PACKAGES = ("model_layer", "model_loss", "model_tool", "model_loader")
FILES = ("model_utility.py", "model_option.py")


def stage(src="/root/reference", verbose=True):
    """Returns the staged path, or None when the reference tree is not present (e.g. on the GPU box)."""
    if not os.path.isdir(src):
        return DST if os.path.isdir(DST) else None
    os.makedirs(DST, exist_ok=True)
    for pkg in PACKAGES:
        d = os.path.join(DST, pkg)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(src, pkg), d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(DST, f))
    with open(os.path.join(DST, "STAGED_FROM"), "w") as fh:
        fh.write(src + "\n")
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print(f"staged {n} files from {src} into {DST}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
