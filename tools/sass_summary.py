#!/usr/bin/env python
"""Counts of the Blackwell-only SASS instructions per kernel of the in-tree library (cuobjdump -sass): the proof that the hot
path is tcgen05 / TMA / TMEM code.  python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "online-gnn-learning_b200", "libogl_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "UBLKCP", "ELECT", "REDG", "ATOMG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            cur = re.sub(r"\(.*$", "", cur)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for p in PAT:
                if op.startswith(p):
                    counts[cur][op] += 1
    print("# SASS mnemonic counts per kernel of %s (sm_100a), `cuobjdump -sass`" % os.path.relpath(LIB, ROOT))
    print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2; kind::f16 and kind::tf32 share the mnemonic), UTMALDG / UTMASTG = TMA tensor")
    print("# load / store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc / dealloc, SYNCS = mbarrier")
    for k, c in counts.items():
        if c:
            print("%-60s %s" % (k[:60], "  ".join("%s x%d" % kv for kv in sorted(c.items()))))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("TOTAL".ljust(60), "  ".join("%s x%d" % kv for kv in sorted(tot.items())))


if __name__ == "__main__":
    sys.exit(main())
