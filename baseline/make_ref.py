"""Populate baseline/_ref/ with the UNMODIFIED reference (python sources only) so that it travels to the GPU box.

    python baseline/make_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

The reference has no setup.py / pyproject, so `pip install --target baseline/_ref /root/reference` has nothing to build;
this is the equivalent "install": a byte-for-byte copy of script/{models,feature,dm,utils} and dataset_loaders/.
baseline/_ref/ is git-ignored (never part of the history) but not gpurun-ignored.  Only bench.py --impl reference and
the drop-in tests (tests/test_reference_dropin_gpu.py) import it, through baseline/ref_shims.py."""
import filecmp
import os
import shutil
import sys

SRC = os.environ.get("DFB_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
TREES = ["script/models", "script/feature", "script/dm", "script/utils", "dataset_loaders"]


def main():
    if not os.path.isdir(SRC):
        print(f"make_ref: {SRC} not present (GPU box): keeping {DST} as shipped")
        return 0
    n = 0
    for t in TREES:
        for root, _, files in os.walk(os.path.join(SRC, t)):
            for f in files:
                if not f.endswith(".py"):
                    continue
                s = os.path.join(root, f)
                d = os.path.join(DST, os.path.relpath(s, SRC))
                os.makedirs(os.path.dirname(d), exist_ok=True)
                if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                    shutil.copyfile(s, d)
                n += 1
    with open(os.path.join(DST, "SOURCE.txt"), "w") as f:
        f.write(f"unmodified copy of {SRC} ({', '.join(TREES)}; *.py only), made by baseline/make_ref.py\n")
    print(f"make_ref: {n} files under {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
