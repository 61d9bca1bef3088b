#!/usr/bin/env python
"""SASS opcode summary of the shipped libmmee.so (evidence that the hot kernels are tcgen05 / TMEM / TMA code):

    python profiles/sass_summary.py > profiles/r2_sass_opcodes.txt

Runs `cuobjdump -sass` on the in-tree library and counts, per kernel, the mnemonics B200_PROFILING.md names:
UTCHMMA (tcgen05.mma, .2CTA = cta_group::2), UTMALDG (TMA tensor load), UTMAPF (TMA L2 prefetch), LDTM / STTM
(tcgen05.ld / .st), UTCBAR (tcgen05.commit), SYNCS (mbarrier), FFMA2 / FADD2 (packed fp32), MUFU.EX2, and — must be
zero — HMMA / WGMMA fallbacks inside the GEMM and attention kernels."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multi-modal-early-exit_b200", "mmee", "libmmee.so")
OPS = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "SYNCS", "FFMA2", "FADD2", "MUFU.EX2",
       "HMMA", "WGMMA", "LDG", "STG", "LDS", "STS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.split("\n")
    counts = collections.OrderedDict()
    cur, i = None, 0
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*", "", names[i]).replace("void ", "").replace("mmee::", "")
            i += 1
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["_total"] += 1
        for o in OPS:
            if o == "UTCHMMA":
                if op.startswith("UTCHMMA") and ".2CTA" not in op:
                    counts[cur][o] += 1
            elif o == "HMMA":
                if op.startswith("HMMA"):
                    counts[cur][o] += 1
            elif op.startswith(o) or (o in op and o.startswith("UTCHMMA.")):
                counts[cur][o] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}  ({len(counts)} kernels)")
    print("kernel".ljust(44) + "".join(o.rjust(13) for o in ["instr"] + OPS))
    tot = collections.Counter()
    for k, c in counts.items():
        print(k[:43].ljust(44) + str(c["_total"]).rjust(13) + "".join(str(c[o]).rjust(13) for o in OPS))
        tot.update(c)
    print("TOTAL".ljust(44) + str(tot["_total"]).rjust(13) + "".join(str(tot[o]).rjust(13) for o in OPS))
    bad = [k for k, c in counts.items() if (k.startswith("gemm_tc") or k.startswith("attention")) and (c["HMMA"] or c["WGMMA"])]
    print("# mma.sync / wgmma fallbacks inside the GEMM / attention kernels:", bad or "none")


if __name__ == "__main__":
    sys.exit(main())
