"""Per-kernel SASS instruction counts of libfreepose_b200.so (what proves a Blackwell-native kernel: B200_PROFILING.md).
    python profiles/sass_summary.py > profiles/sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "freepose_b200" / "libfreepose_b200.so"
MNEMONICS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "MUFU.EX2", "MUFU.RCP",
             "FFMA2", "ATOM", "ATOMS", "RED", "LDGSTS", "STL", "LDL"]

sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
out = []
for blk in sass.split("Function : ")[1:]:
    name = blk.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(anonymous namespace\)::", "", dem).split("(")[0].replace("void fp::", "").replace("fp::", "")
    ops = collections.Counter()
    n = 0
    for line in blk.splitlines():
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        n += 1
        op = m.group(1)
        for key in MNEMONICS:
            if key == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    ops[key] += 1
            elif key in ("HMMA",):
                if op.startswith("HMMA"):
                    ops[key] += 1
            elif op.startswith(key):
                ops[key] += 1
    out.append((dem, n, ops))
print("# cuobjdump -sass freepose_b200/libfreepose_b200.so (sm_100a), instructions per kernel by mnemonic")
print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA load/store, HMMA = legacy mma.sync")
hdr = ["kernel", "instr"] + MNEMONICS
print(" | ".join(hdr))
for dem, n, ops in sorted(out):
    print(" | ".join([dem, str(n)] + [str(ops.get(k, 0)) for k in MNEMONICS]))
tot = collections.Counter()
for _, _, ops in out:
    tot.update(ops)
print("# totals: " + ", ".join(f"{k} {tot.get(k, 0)}" for k in MNEMONICS))
