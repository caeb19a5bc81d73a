#!/usr/bin/env python
"""SASS digest of the built CUDA library: per kernel, how many global loads are 128-bit (LDG.E.128), how many asynchronous /
bulk copies (LDGSTS, UBLKCP), global atomics / reductions (ATOMG, RED), shared-memory accesses (LDS / STS), warp shuffles and
votes, plus registers and static shared memory from `cuobjdump -res-usage`.
usage: python tools/sass_digest.py [aletsch_b200/libaletsch_gpu.so] > profiles/rNN_sass_digest.md"""
import collections
import re
import subprocess
import sys

PAT = [("LD generic", r"\bLD\.E"), ("LDG", r"\bLDG\."), ("LDG.E.128", r"\bLDG\.E\.(?:\w+\.)*128"), ("LDG.E.64", r"\bLDG\.E\.(?:\w+\.)*64"), ("STG", r"\bSTG\."),
       ("STG.E.128", r"\bSTG\.E\.(?:\w+\.)*128"), ("LDGSTS", r"\bLDGSTS"), ("UBLKCP", r"\bUBLKCP"), ("ATOMG", r"\bATOMG"), ("RED", r"\bRED\."),
       ("LDS", r"\bLDS"), ("STS", r"\bSTS"), ("ATOMS", r"\bATOMS"), ("SHFL", r"\bSHFL"), ("VOTE/MATCH", r"\b(?:VOTE|MATCH)"), ("BAR", r"\bBAR\.")]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else "aletsch_b200/libaletsch_gpu.so"
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            r = re.search(r"REG:(\d+)", line)
            sh = re.search(r"SHARED:(\d+)", line)
            usage[cur] = (int(r.group(1)) if r else 0, int(sh.group(1)) if sh else 0)
            cur = None
    acc = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = acc.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None or "/*" not in line:
            continue
        for name, pat in PAT:
            if re.search(pat, line):
                cur[name] += 1
    arch = re.findall(r"\.(sm_\w+)\.cubin", subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout)
    names = list(acc.keys())
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    pretty = {n: (d.split("(")[0].replace("agpu::", "") if d else n) for n, d in zip(names, dem)}
    print("# SASS digest of %s" % so)
    print()
    print("`cuobjdump -sass` / `-res-usage`; ELF images: %s.  Counts are static instruction counts per kernel." % (", ".join(sorted(set(arch))) or "?"))
    print()
    tot = collections.Counter()
    print("| kernel | regs | smem B | " + " | ".join(n for n, _ in PAT) + " |")
    print("|---|---:|---:|" + "---:|" * len(PAT))
    for k, c in acc.items():
        short = pretty.get(k, k)
        r, sh = usage.get(k, (0, 0))
        print("| %s | %d | %d | %s |" % (short, r, sh, " | ".join(str(c[n]) for n, _ in PAT)))
        tot.update(c)
    print("| **all** | | | %s |" % " | ".join(str(tot[n]) for n, _ in PAT))


if __name__ == "__main__":
    main()
