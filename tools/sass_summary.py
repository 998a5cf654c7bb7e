#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libklerg_b200.so (cuobjdump -sass): which kernels carry packed FP32 (FFMA2 /
FADD2 / FMUL2), MUFU.EX2, TMA bulk copies (UBLKCP), tcgen05 MMAs (UTCHMMA / UTCQMMA ...), TMEM stores / loads
(STTM / LDTM), mbarrier waits (SYNCS), programmatic dependent launch (ACQBULK = griddepcontrol.wait, PREEXIT =
griddepcontrol.launch_dependents) etc.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "embodied-active-learning-vision_b200", "libklerg_b200.so")
WATCH = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "MUFU.EX2", "MUFU.LG2", "MUFU.RCP", "DADD", "DFMA", "UBLKCP",
         "UTCHMMA", "UTCQMMA", "UTCBAR", "STTM", "LDTM", "SYNCS", "LDS", "STS", "LDG", "STG", "LD.E", "ST.E", "BAR", "ACQBULK", "PREEXIT",
         "LDL", "STL", "CCTL", "MEMBAR", "ATOM", "RED", "SHFL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[cur][w] += 1
                    break
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.basename(LIB)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a); only kernels with > 200 instructions")
    print(f"# columns: total | " + " ".join(WATCH))
    for (k, c), nm in zip(kernels.items(), names):
        if c["_total"] < 200 or "_emu_" in nm:
            continue
        nm = re.sub(r"\(klerg::.*", "", nm).replace("void klerg::", "")
        row = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
        print(f"{nm[:90]:90s} {c['_total']:6d} | {row}")


if __name__ == "__main__":
    main()
