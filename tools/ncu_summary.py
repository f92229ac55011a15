#!/usr/bin/env python3
"""Condenses one `ncu --set full` capture of the search kernel (the CSV exports of tools/profile*.sh:
`--page raw --csv` and `--page source --csv`) into the text summary kept under profiles/.

    python tools/ncu_summary.py RAW.csv SOURCE.csv CANDIDATES_PER_LAUNCH > profiles/<name>_ncu_summary.txt
"""
import collections
import csv
import re
import sys

KEEP = [
    (r"^gpu__time_duration\.sum$", "kernel duration"),
    (r"^launch__grid_size$", "grid"), (r"^launch__block_size$", "block"), (r"^launch__registers_per_thread$", "registers / thread"),
    (r"^launch__shared_mem_per_block_dynamic$", "dynamic shared memory / block"),
    (r"^smsp__inst_executed\.sum$", "warp instructions executed"),
    (r"^sm__inst_executed\.avg\.per_cycle_active$", "IPC (per SM, active)"),
    (r"^sm__inst_issued\.avg\.pct_of_peak_sustained_active$", "issue slots busy"),
    (r"^sm__inst_executed_pipe_(alu|fma|xu|lsu|adu|cbu|uniform)\.avg\.pct_of_peak_sustained_active$", None),
    (r"^smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio$", None),
    (r"^dram__bytes_read\.sum$", "DRAM bytes read"), (r"^dram__bytes_write\.sum$", "DRAM bytes written"),
    (r"^lts__t_bytes\.sum$", "L2 bytes"),
    (r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$", "shared-memory wavefronts"),
    (r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$", "shared-memory bank conflicts"),
    (r"^smsp__sass_inst_executed_op_local_(ld|st)\.sum$", None),
    (r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", "achieved occupancy"),
]


def opcode(text):
    t = text.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0] if t else "?"


def main():
    raw, src, ncand = sys.argv[1], sys.argv[2], float(sys.argv[3])
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    print("kernel:", d.get("Kernel Name", ("", "?"))[1])
    print("candidates per launch: %d" % ncand)
    for pat, label in KEEP:
        for h in hdr:
            if re.search(pat, h):
                u, v = d[h]
                try:
                    if label is None and float(v) < 0.02:
                        continue
                except ValueError:
                    pass
                name = label or h.replace("smsp__average_warps_issue_stalled_", "stall: ").replace("_per_issue_active.ratio", " (warps per issue)") \
                                  .replace("sm__inst_executed_pipe_", "pipe ").replace(".avg.pct_of_peak_sustained_active", " busy")
                print("  %-58s %s %s" % (name, v, u))
    try:
        total = float(d["smsp__inst_executed.sum"][1])
        dur_ms = float(d["gpu__time_duration.sum"][1]) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(d["gpu__time_duration.sum"][0], 1.0)
        dram = sum(float(d[k][1]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(d[k][0], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        print("derived:")
        print("  warp instructions per candidate                            %.1f" % (total / ncand))
        print("  candidates per second (under ncu: not a benchmark value)   %.3e" % (ncand / (dur_ms * 1e-3)))
        print("  DRAM traffic per launch (read + write)                     %.0f bytes" % dram)
    except (KeyError, ValueError):
        pass
    rows = list(csv.reader(open(src)))
    hdr, data = rows[1], rows[2:]
    ia, isrc, iex, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = int(data[0][ia], 16)
    tot, samp, lines = collections.Counter(), collections.Counter(), []
    for r in data:
        ex, s = int(r[iex]), int(r[isamp])
        o = opcode(r[isrc])
        tot[o] += ex
        samp[o] += s
        lines.append((s, int(r[ia], 16) - base, ex, r[isrc].strip()))
    S = sum(samp.values()) or 1
    print("executed instruction mix (warp instructions per candidate, share of the stall samples):")
    for o, c in tot.most_common(16):
        print("  %-10s %8.1f   %5.1f %%" % (o, c / ncand, 100.0 * samp[o] / S))
    print("hottest SASS instructions (stall samples, offset in the kernel, executions per candidate):")
    for s, a, ex, text in sorted(lines, reverse=True)[:24]:
        print("  %6d  %#07x  %7.2f  %s" % (s, a, ex / ncand, text))


if __name__ == "__main__":
    main()
