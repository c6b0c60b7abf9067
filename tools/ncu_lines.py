#!/usr/bin/env python
"""Executed instructions and stall samples of one kernel of an `ncu --set full --import-source on`
capture, attributed to SOURCE LINES.

ncu's source page (CSV) lists the kernel's SASS in program order with "Instructions Executed"
and "# Samples" per instruction, but no line numbers; `nvdisasm --print-line-info` on the cubin
of the same build lists the same SASS in the same order with `//## File "...", line N` markers.
The two are matched by position (and checked by opcode).

  cuobjdump -xelf all psc_b200/csrc/build/push_exact.o           # -> push.sm_100a.cubin
  python tools/ncu_lines.py capture.ncu-rep push.sm_100a.cubin k_push_leanILi0ELi1ELb1ELb1ELi1 [--top 40]

The kernel is selected by a substring of its mangled name.  Used for DESIGN.md 3.2d / 8 (the
per-line instruction diff between k_push_lean and k_push_lean_pull) and the collision kernel."""
import argparse
import collections
import csv
import io
import os
import re
import subprocess


def disasm_lines(cubin, name):
    txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(txt) if l.startswith(".text.") and name in l)
    end = start + 1
    while end < len(txt) and not txt[end].startswith("//--------------------- .text."):
        end += 1
    cur, out = None, []
    for l in txt[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
        if m:
            out.append((cur, m.group(1).strip()))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("cubin")
    ap.add_argument("kernel", help="substring of the mangled kernel name")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--src", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "psc_b200", "csrc"))
    args = ap.parse_args()
    page = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(page)))
    h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr = rows[h]
    i_src, i_ex, i_s = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ncu = [(r[i_src].strip(), int(r[i_ex]), int(r[i_s])) for r in rows[h + 1:] if len(r) > i_ex]
    dis = disasm_lines(args.cubin, args.kernel)
    n = min(len(ncu), len(dis))
    if len(ncu) != len(dis):
        print("warning: %d instructions in the capture, %d in the cubin (different builds?)" % (len(ncu), len(dis)))
    ex, sm = collections.Counter(), collections.Counter()
    mism = 0
    for k in range(n):
        a = ncu[k][0].split()
        b = dis[k][1].split()
        oa = a[1] if a and a[0].startswith("@") and len(a) > 1 else (a[0] if a else "")
        ob = b[1] if b and b[0].startswith("@") and len(b) > 1 else (b[0] if b else "")
        mism += oa.split(".")[0] != ob.split(".")[0]
        ex[dis[k][0]] += ncu[k][1]
        sm[dis[k][0]] += ncu[k][2]
    tot, ts = sum(ex.values()), max(1, sum(sm.values()))
    print("kernel %s: %.3e warp instructions, %d stall samples, %d opcode mismatches" % (args.kernel, tot, ts, mism))
    cache = {}
    for key, v in ex.most_common(args.top):
        text = ""
        if key:
            path = os.path.join(args.src, key[0])
            if os.path.exists(path):
                cache.setdefault(path, open(path).read().split("\n"))
                text = cache[path][key[1] - 1].strip()[:80]
        print("%-28s %5d  %.2e %5.1f%%  samples %5.1f%%  | %s" % (key[0] if key else "?", key[1] if key else 0, v,
                                                                  100. * v / tot, 100. * sm[key] / ts, text))


if __name__ == "__main__":
    main()
