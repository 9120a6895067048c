#!/usr/bin/env python
"""SASS opcode histogram of the shipped library: the Blackwell-native instruction evidence (tcgen05 = UTCHMMA / UTCBAR, TMA = UTMALDG /
UTMASTG, TMEM loads = LDTM), per kernel family.

    python tools/sass_histogram.py geo-trax_b200/libgeotrax_b200.so > profiles/sass_histogram_r2.txt
"""
import collections
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "UTCATOM", "SYNCS", "MUFU", "POPC", "LDS", "STS", "LDG", "STG", "ELECT", "BAR"]


def main():
    so = sys.argv[1]
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    fam = None
    per = collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fam = name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
            per.setdefault(fam, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
        if m and fam:
            op = m.group(1)
            per[fam]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    per[fam][op if k in ("UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM") else k] += 1
    print(f"# SASS opcode histogram of {so} (cuobjdump -sass; sm_100a)")
    tot = collections.Counter()
    for fam, c in per.items():
        if c["_total"] == 0:
            continue
        print(f"\n{fam}: {c['_total']} instructions")
        for k, v in sorted(c.items()):
            if k != "_total":
                print(f"    {k:32s} {v}")
                tot[k] += v
    print("\n# whole library")
    for k, v in sorted(tot.items()):
        print(f"    {k:32s} {v}")


if __name__ == "__main__":
    main()
