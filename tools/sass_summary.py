#!/usr/bin/env python
"""SASS evidence for the shipped code object: per kernel of libautofunc_b200.so (sm_100a), the register count, shared
memory and a histogram of the instructions that prove the data path -- DMMA.8x8x4 (FP64 tensor core), UTMALDG (TMA tensor
loads), SYNCS (mbarrier), UCGABAR / ST.E...cluster forms (thread-block clusters, distributed shared memory), REDG / ATOMG
(global float64 atomics), BAR, SHFL -- plus the total instruction count.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "autofunc_b200", "libautofunc_b200.so")
KEYS = [("DMMA", r"^DMMA"), ("UTMALDG", r"^UTMALDG"), ("SYNCS", r"^SYNCS"), ("cluster/DSMEM", r"UCGABAR|CGABAR|\.CLUSTER|MAPA|STS\.\w*CLUSTER|ST\.\w*\.CLUSTER"),
        ("REDG/ATOMG f64", r"^(REDG|ATOMG|RED|ATOM)\b.*F64|^(REDG|ATOMG)"), ("BAR", r"^BAR"), ("SHFL", r"^SHFL"), ("LDS", r"^LDS"), ("LDG", r"^LDG"),
        ("STG", r"^STG"), ("DFMA/DADD/DMUL", r"^(DFMA|DADD|DMUL)"), ("MUFU", r"^MUFU"), ("BPT.TRAP", r"^BPT")]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out)) if len(out) == len(names) else {n: n for n in names}


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
        usage[m.group(1)] = dict(reg=int(m.group(2)), stack=int(m.group(3)), shared=int(m.group(4)), local=int(m.group(5)))
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m and cur is not None:
            op, rest = m.group(1), m.group(2)
            cur["total"] += 1
            for key, pat in KEYS:
                if re.search(pat, op) or (key == "cluster/DSMEM" and re.search(pat, op + rest)):
                    cur[key] += 1
    names = demangle(list(kernels))
    print(f"# SASS summary of autofunc_b200/libautofunc_b200.so ({', '.join(arch)}): {len(kernels)} kernels")
    print("# kernel | registers | static smem B | local B | total instr | " + " | ".join(k for k, _ in KEYS))
    tot = collections.Counter()
    for k, c in sorted(kernels.items(), key=lambda kv: -kv[1]["DMMA"]):
        u = usage.get(k, {})
        short = re.sub(r"\((int|bool)\)", "", names.get(k, k)).replace("(anonymous namespace)::", "")
        short = re.sub(r"\(.*", "", short).replace("void ", "").replace("af::", "")
        print(f"{short} | {u.get('reg', '?')} | {u.get('shared', '?')} | {u.get('local', '?')} | {c['total']} | " + " | ".join(str(c[key]) for key, _ in KEYS))
        tot.update(c)
    print("# all kernels | - | - | - | " + f"{tot['total']} | " + " | ".join(str(tot[key]) for key, _ in KEYS))
    return 0


if __name__ == "__main__":
    sys.exit(main())
