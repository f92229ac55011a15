#!/usr/bin/env python3
"""Per-source-line view of one ncu capture of a search kernel: joins the SASS page of the capture
(`ncu -i X.ncu-rep --page source --csv`, one row per instruction with its executed count and stall
samples) with the line table of the same kernel in the built object (`nvdisasm -g -c`, needs -lineinfo).

    python tools/ncu_by_line.py SOURCE.csv OBJECT.o 'EvalCfgTILi2ELi32ELi1ELi139810ELi640' CANDIDATES [--top N]

Prints warp instructions per candidate and stall-sample share per (file, line), inlined call sites
attributed to the innermost line, and a per-file-region summary.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def line_table(obj, needle):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cub = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.splitlines()
    table, cur, inside = {}, None, False
    for l in dis:
        if l.startswith(".text."):
            inside = needle in l
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    src, obj, needle, ncand = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    table = line_table(obj, needle)
    rows = list(csv.reader(open(src)))
    hdr, data = rows[1], rows[2:]
    ia, iex, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = int(data[0][ia], 16)
    per_line, per_file = collections.Counter(), collections.Counter()
    samp_line = collections.Counter()
    total = samples = 0
    for r in data:
        off = int(r[ia], 16) - base
        ex, s = int(r[iex]), int(r[isamp])
        loc = table.get(off, (None, ""))[0] or ("?", 0)
        per_line[loc] += ex
        samp_line[loc] += s
        total += ex
        samples += s
    print("kernel filter: %s   instructions matched to lines: %d of %d SASS rows" % (needle, sum(1 for r in data if (int(r[ia], 16) - base) in table), len(data)))
    print("warp instructions per candidate: %.1f" % (total / ncand))
    print("%-28s %10s %8s" % ("file:line", "inst/cand", "stall %"))
    for loc, ex in per_line.most_common(top):
        print("%-28s %10.1f %7.1f%%" % ("%s:%d" % loc, ex / ncand, 100.0 * samp_line[loc] / max(samples, 1)))


if __name__ == "__main__":
    main()
