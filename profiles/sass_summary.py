#!/usr/bin/env python
"""Per-kernel SASS evidence of libespnet_b200.so: counts of the Blackwell mnemonics that prove tcgen05 / TMEM / TMA
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG, cp.async.bulk -> UBLKCP;
legacy mma.sync would show as HMMA).  usage: python profiles/sass_summary.py > profiles/sass_summary.txt  (no GPU needed)"""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "glomeruli_segmentation_b200", "csrc", "libespnet_b200.so")
MNEM = ["UTCHMMA", "A_KEEP", "A_REUSE", "LDTM", "UTMALDG", "UBLKCP", "UTMACCTL", "SYNCS", "ACQBULK", "PREEXIT", "HMMA", "FFMA2", "FFMA", "LDG", "STG", "LDS", "STS", "ATOM", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {k: 0 for k in MNEM}
            kernels[cur]["instr"] = 0
            continue
        if cur is None or not re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            continue
        kernels[cur]["instr"] += 1
        body = line.split("*/", 1)[1]
        for k in MNEM:
            if k in ("A_KEEP", "A_REUSE"):
                kernels[cur][k] += body.count("." + k)
            elif re.search(r"\b%s(\.|\b)" % k, body) and not (k == "FFMA" and "FFMA2" in body) and not (k == "HMMA" and "UTCHMMA" in body):
                kernels[cur][k] += 1
    dem = subprocess.run(["cu++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonic counts per kernel of %s (architectures in the fatbin: %s)" % (os.path.relpath(SO, ROOT), ", ".join(arch)))
    print("# UTCHMMA = tcgen05.mma (.A_KEEP / .A_REUSE = A-collector pairs), LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk,")
    print("# SYNCS = mbarrier, ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents (programmatic dependent launch), HMMA = legacy mma.sync (none expected), FFMA2 = packed fp32x2 FMA")
    cols = ["instr"] + MNEM
    print("%-86s " % "kernel" + " ".join("%8s" % c for c in cols))
    tot = {c: 0 for c in cols}
    for (mangled, c), name in zip(kernels.items(), dem):
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\((int|bool)\)", "", name)
        name = re.sub(r"\(.*$", "", name).replace("espnet::", "")
        print("%-86s " % name[:86] + " ".join("%8d" % c[k] for k in cols))
        for k in cols:
            tot[k] += c[k]
    print("%-86s " % ("TOTAL (%d kernels)" % len(kernels)) + " ".join("%8d" % tot[k] for k in cols))


if __name__ == "__main__":
    main()
