"""SASS / ptxas summary of libudt_b200.so (no GPU needed): per kernel, the ptxas resource line (registers, barriers, stack /
spill) from the build log and the count of the instructions that prove the Blackwell-native paths (B200_PROFILING.md):
UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor load / store),
UBLKCP (bulk copy), SYNCS (mbarrier), MUFU, and legacy HMMA (must be 0).
usage: python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "udifftext_b200", "libudt_b200.so")
LOG = os.path.join(ROOT, "udifftext_b200", "build", "nvcc.log")
PATS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "MUFU.EX2", "MUFU", "HMMA."]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", o) for o in out]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {p: 0 for p in PATS}
            kernels[cur]["instructions"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(1)
        kernels[cur]["instructions"] += 1
        op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
        for p in PATS:
            if op.startswith(p):
                kernels[cur][p] += 1
    for k in kernels.values():            # the generic patterns include their specialisations: make them exclusive
        k["UTCHMMA"] -= k["UTCHMMA.2CTA"]
        k["MUFU"] -= k["MUFU.EX2"]
    res = {}
    if os.path.exists(LOG):
        txt = open(LOG).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n(?:ptxas info\s+: Function properties for \S+\n\s+(.*)\n)?ptxas info\s+: Used (.*)", txt):
            res[m.group(1)] = (m.group(3), m.group(2) or "")
    names = list(kernels)
    pretty = demangle(names)
    tot = {p: sum(k[p] for k in kernels.values()) for p in PATS}
    print("libudt_b200.so — SASS summary (cuobjdump -sass, sm_100a) and ptxas -v resources")
    print("totals: " + ", ".join(f"{p} {tot[p]}" for p in PATS) + f"; kernels {len(kernels)}")
    print()
    for n, pn in sorted(zip(names, pretty), key=lambda t: -kernels[t[0]]["instructions"]):
        k = kernels[n]
        marks = ", ".join(f"{p} {k[p]}" for p in PATS if k[p])
        print(f"{pn[:150]}")
        print(f"    {k['instructions']} instructions; {marks or 'no tensor / TMA instructions'}")
        if n in res:
            print(f"    ptxas: {res[n][0]}" + (f"; {res[n][1]}" if res[n][1] else ""))
    assert tot["HMMA."] == 0, "legacy mma.sync found"


if __name__ == "__main__":
    main()
