#!/usr/bin/env python
"""Stall-reason samples of an ncu capture attributed to source lines (innermost line of the -lineinfo chain).

    python profiles/ncu_stalls.py <report.ncu-rep> <kernel mangled-name substring> [top_n]
"""
import collections
import csv
import subprocess
import sys

import ncu_lines

REASONS = ("stall_long_sb", "stall_barrier", "stall_wait", "stall_short_sb", "stall_branch_resolving",
           "stall_not_selected", "stall_no_inst", "stall_math", "stall_mio", "stall_lg", "stall_dispatch",
           "stall_selected")


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    chains = ncu_lines.disasm_chains(ksub)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    col = {h: i for i, h in enumerate(rows[hi])}
    per = {r: collections.Counter() for r in REASONS}
    tot = collections.Counter()
    base = None
    for r in rows[hi + 1:]:
        if len(r) < len(rows[hi]) or not r[0].startswith("0x"):
            continue
        addr = int(r[0], 16)
        base = addr if base is None else base
        chain, text = chains.get(addr - base, ([("?", 0)], r[1]))
        key = "%s:%d" % (chain[0] if chain else ("?", 0))
        outer = "%s:%d" % (chain[-1] if chain else ("?", 0))
        op = (text.split()[1] if text.startswith("@") else text.split()[0]).split(".")[0]
        for reason in REASONS:
            v = int(r[col[reason]])
            if v:
                per[reason][(key, outer, op)] += v
                tot[reason] += v
    total = sum(tot.values())
    for reason in sorted(REASONS, key=lambda x: -tot[x]):
        if tot[reason] == 0:
            continue
        print("== %s: %d samples (%.1f%%)" % (reason, tot[reason], 100.0 * tot[reason] / total))
        for (key, outer, op), v in per[reason].most_common(top):
            print("   %-24s in %-22s %-8s %6d" % (key, outer, op, v))


if __name__ == "__main__":
    main()
