#!/usr/bin/env python
"""Attribute an ncu capture's per-SASS-instruction counters to CUDA source lines.

    python profiles/ncu_lines.py <report.ncu-rep> <kernel mangled-name substring> [top_n]

Joins `ncu --page source --csv` (per SASS instruction: executed warp instructions, thread instructions,
stall samples, shared-memory wavefronts) with `nvdisasm --print-line-info-inline` of the cubins in
csrc/libtetris_b200.so (built with -lineinfo; TB_SO_PATH overrides the library), by instruction offset.  Prints
totals per innermost source line, per kernel-body line and per opcode, and the shared-memory bank-conflict table:
wavefronts, ideal wavefronts and excessive (= conflict replay) wavefronts per source line.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

DEPTH = int(os.environ.get("NCU_LINES_DEPTH", "1"))   # 0: the __global__ function's line; 1: one inlining level below
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.environ.get("TB_SO_PATH") or os.path.join(ROOT, "tetris_b200", "csrc", "libtetris_b200.so")


def disasm_chains(kernel_sub):
    """offset -> [(file, line) innermost .. outermost]"""
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", SO], cwd=tmp, stdout=subprocess.DEVNULL)
    txt = []                                   # one cubin per translation unit (board shape, distinct names): scan them all
    for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin") and "sm_100" in f):
        txt += subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", os.path.join(tmp, cubin)],
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
    shm = collections.defaultdict(lambda: [0, 0, 0, 0])          # wavefronts, ideal, excessive, instructions
    shm_tot = [0, 0, 0]
    tot = [0, 0, 0]

    def num(r, name):
        i = col.get(name)
        try:
            return int(float(r[i])) if i is not None and r[i] not in ("", "-") else 0
        except ValueError:
            return 0
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
        if DEPTH and len(chain) > DEPTH:
            outer = chain[-1 - DEPTH]          # the line inside the body function the kernel wraps (phase attribution)
        for d, k in ((by_line, inner), (by_outer, outer)):
            d[k][0] += inst; d[k][1] += thr; d[k][2] += smp
        op = text.split()[0] if not text.startswith("@") else text.split()[1]
        op = op.split(".")[0]
        by_op[op][0] += inst; by_op[op][1] += thr; by_op[op][2] += smp
        tot[0] += inst; tot[1] += thr; tot[2] += smp
        wf, ideal, exc = num(r, "L1 Wavefronts Shared"), num(r, "L1 Wavefronts Shared Ideal"), num(r, "L1 Wavefronts Shared Excessive")
        if wf:
            key = (inner, outer, op)
            shm[key][0] += wf; shm[key][1] += ideal; shm[key][2] += exc; shm[key][3] += inst
            shm_tot[0] += wf; shm_tot[1] += ideal; shm_tot[2] += exc
    print("total warp-inst %d  thread-inst %d  samples %d  (avg active threads %.1f)" % (tot[0], tot[1], tot[2], tot[1] / max(tot[0], 1)))
    for title, d in (("innermost source line", by_line), ("outermost (kernel-body) line", by_outer), ("opcode", by_op)):
        print("\n== by %s: warp-inst%%  thread-inst%%  samples%%  avg-threads" % title)
        for k, v in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
            name = "%s:%d" % k if isinstance(k, tuple) else k
            print("%-28s %6.2f %6.2f %6.2f  %5.1f" % (name, 100.0 * v[0] / tot[0], 100.0 * v[1] / max(tot[1], 1),
                                                     100.0 * v[2] / max(tot[2], 1), v[1] / max(v[0], 1)))
    print_shared(shm, shm_tot, top)


def print_shared(shm, shm_tot, top):
    print("\n== shared memory: %d wavefronts, %d ideal, %d excessive (bank-conflict replays) = %.1f %% of wavefronts"
          % (shm_tot[0], shm_tot[1], shm_tot[2], 100.0 * shm_tot[2] / max(shm_tot[0], 1)))
    print("   innermost line <- kernel-body line, opcode: excessive%%-of-all-excessive  wavefronts  ideal  excessive  wavefronts/inst")
    for (inner, outer, op), v in sorted(shm.items(), key=lambda kv: -kv[1][2])[:top]:
        print("%-22s <- %-22s %-5s %6.2f %11d %11d %11d  %5.2f" % ("%s:%d" % inner, "%s:%d" % outer, op,
              100.0 * v[2] / max(shm_tot[2], 1), v[0], v[1], v[2], v[0] / max(v[3], 1)))


if __name__ == "__main__":
    main()
