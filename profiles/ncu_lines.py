#!/usr/bin/env python
"""Attribute an ncu capture's per-SASS-instruction counters to CUDA source lines.

    python profiles/ncu_lines.py <report.ncu-rep> <kernel mangled-name substring> [top_n]

Joins `ncu --page source --csv` (per SASS instruction: executed warp instructions, thread instructions,
stall samples) with `nvdisasm --print-line-info-inline` of the cubin in csrc/libtetris_b200.so (built with
-lineinfo), by instruction offset.  Prints totals per innermost source line and per inlined callee.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tetris_b200", "csrc", "libtetris_b200.so")


def disasm(kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", SO], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", os.path.join(tmp, cubin)],
                         capture_output=True, text=True).stdout.splitlines()
    out, cur, active = {}, [], False
    for ln in txt:
        if ln.startswith(".text.") and ln.endswith(":"):
            active = kernel_sub in ln
            cur = []
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            cur.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            off = int(m.group(1), 16)
            # `cur` = chain innermost ... outermost for this instruction group; it persists until the next //## block
            out[off] = list(cur)
            pending = True
        elif ln.strip().startswith(".L_") or not ln.strip():
            pass
        # a new //## group after an instruction starts a fresh chain
        if m is None:
            continue
        cur_after = cur
        cur = cur_after
    return out


def disasm_chains(kernel_sub):
    """offset -> [(file, line) innermost .. outermost]"""
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", SO], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", os.path.join(tmp, cubin)],
                         capture_output=True, text=True).stdout.splitlines()
    out, chain, fresh, active = {}, [], True, False
    for ln in txt:
        if ln.startswith(".text.") and ln.endswith(":"):
            active = kernel_sub in ln
            chain, fresh = [], True
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            if fresh:
                chain, fresh = [], False
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out[int(m.group(1), 16)] = (list(chain), m.group(2).strip())
            fresh = True
    return out


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    chains = disasm_chains(ksub)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    base = None
    by_line = collections.defaultdict(lambda: [0, 0, 0])
    by_outer = collections.defaultdict(lambda: [0, 0, 0])
    by_op = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        addr = int(r[0], 16)
        if base is None:
            base = addr
        off = addr - base
        inst = int(r[col["Instructions Executed"]]); thr = int(r[col["Thread Instructions Executed"]])
        smp = int(r[col["# Samples"]])
        chain, text = chains.get(off, ([("?", 0)], r[1]))
        inner = chain[0] if chain else ("?", 0)
        outer = chain[-1] if chain else ("?", 0)
        for d, k in ((by_line, inner), (by_outer, outer)):
            d[k][0] += inst; d[k][1] += thr; d[k][2] += smp
        op = text.split()[0] if not text.startswith("@") else text.split()[1]
        op = op.split(".")[0]
        by_op[op][0] += inst; by_op[op][1] += thr; by_op[op][2] += smp
        tot[0] += inst; tot[1] += thr; tot[2] += smp
    print("total warp-inst %d  thread-inst %d  samples %d  (avg active threads %.1f)" % (tot[0], tot[1], tot[2], tot[1] / max(tot[0], 1)))
    for title, d in (("innermost source line", by_line), ("outermost (kernel-body) line", by_outer), ("opcode", by_op)):
        print("\n== by %s: warp-inst%%  thread-inst%%  samples%%  avg-threads" % title)
        for k, v in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
            name = "%s:%d" % k if isinstance(k, tuple) else k
            print("%-28s %6.2f %6.2f %6.2f  %5.1f" % (name, 100.0 * v[0] / tot[0], 100.0 * v[1] / max(tot[1], 1),
                                                     100.0 * v[2] / max(tot[2], 1), v[1] / max(v[0], 1)))


if __name__ == "__main__":
    main()
