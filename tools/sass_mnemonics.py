"""SASS mnemonic counts per kernel of the built library (cuobjdump -sass): the evidence that the hot kernels use the
sm_100a tensor-core / TMA / TMEM instructions, programmatic dependent launch and vector atomics.
    python tools/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "dfnet_b200", "libdfnet_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
COLS = [("UTCHMMA", r"\bUTC[HQ]?MMA"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"), ("UTCBAR", r"\bUTCBAR"),
        ("SYNCS.ARRIVE", r"\bSYNCS\.ARRIVE"), ("MEMBAR.ALL.GPU", r"\bMEMBAR\.ALL\.GPU"), ("ACQBULK", r"\bACQBULK"),
        ("PREEXIT", r"\bPREEXIT"), ("REDG.F32x4", r"\bREDG\.E\.ADD\.F32x4")]
print("# SASS mnemonic counts per kernel of dfnet_b200/libdfnet_b200.so that uses tcgen05 / TMA / dependent launch (cuobjdump -sass,")
print("# sm_100a), final round-2 build.  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk,")
print("# UTCBAR = tcgen05.commit, ACQBULK = griddepcontrol.wait, PREEXIT = griddepcontrol.launch_dependents, REDG.F32x4 = 16-byte")
print("# vector atomic add.  MEMBAR.ALL.GPU: only the two cluster barriers at kernel start / end remain (see DESIGN 4.1a).")
print("  ".join(f"{c[0]:>{max(len(c[0]), 5)}s}" for c in COLS) + "  kernel")
parts = re.split(r"\n\s*Function : \S+\n", sass)[1:]
for nm, body in zip(names, parts):
    cnt = [len(re.findall(rx, body)) for _, rx in COLS]
    if cnt[0] + cnt[2] + cnt[7] + cnt[9] == 0:
        continue
    print("  ".join(f"{v:>{max(len(c[0]), 5)}d}" for v, c in zip(cnt, COLS)) + "  " + nm[:110])
