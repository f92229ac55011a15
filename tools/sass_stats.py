#!/usr/bin/env python3
"""Static SASS statistics of one kernel of an object file: instruction count, opcode histogram,
loops (backward branches) with their body sizes and opcode mix.  Used to compare builds of the hot
kernel without a GPU: the evaluator's row loops run a fixed number of times per candidate
(config 3: 3 iterations of the unchecked two-tile loop + 1 of the checked one), so static body
sizes translate into executed instructions per candidate.

    python tools/sass_stats.py kafka_assignment_optimizer_b200/_obj/kao_inst_full_2_3.o \
        'search_persistent_kernel<kao::EvalCfg<2, 3, 3, 3>, 768, false>'
"""
import collections
import re
import subprocess
import sys


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True, check=True).stdout
    cur, res = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            res[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur is not None:
            res[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return res


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    return dict(zip(names, out))


def opcode(text):
    t = text.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0] if t else "?"


def main():
    obj, want = sys.argv[1], sys.argv[2]
    ks = kernels(obj)
    dm = demangle(list(ks))
    hits = [k for k in ks if dm[k].startswith("void " + want + "(") or dm[k].startswith(want + "(")]
    if len(hits) != 1:
        sys.exit("kernel not found or ambiguous: %r among\n  %s" % (want, "\n  ".join(sorted(dm.values()))))
    ins = ks[hits[0]]
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    print("%s\n  %d instructions" % (dm[hits[0]], len(ins)))
    hist = collections.Counter(opcode(t) for _, t in ins)
    print("  " + "  ".join("%s %d" % kv for kv in hist.most_common(14)))
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?`\(\.L_x_\d+\)|\bBRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m and m.group(1):
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], i))
    for lo, hi in sorted(loops):
        body = ins[lo:hi + 1]
        h = collections.Counter(opcode(t) for _, t in body)
        print("  loop %#06x..%#06x  %4d instructions   %s" % (ins[lo][0], ins[hi][0], len(body),
              "  ".join("%s %d" % kv for kv in h.most_common(9))))


if __name__ == "__main__":
    main()
