#!/usr/bin/env python3
"""Prints the variants the tuning probe measured (bench line -> config.evaluator.variants, or the raw
`PROBE {json}` lines of `bench.py --probe-evaluators`), fastest first.

    python tools/tune_report.py BENCH_r01.json
    python bench.py --probe-evaluators | python tools/tune_report.py -
"""
import json
import sys


def main():
    text = sys.stdin.read() if sys.argv[1:] == ["-"] else open(sys.argv[1]).read()
    rows = []
    for line in text.splitlines():
        line = line.strip()
        if line.startswith("PROBE "):
            rows.append(json.loads(line[6:]))
        elif line.startswith("{"):
            try:
                d = json.loads(line)
            except ValueError:
                continue
            ev = (d.get("config") or {}).get("evaluator") or {}
            rows += ev.get("variants", [])
            if ev.get("selected"):
                print("selected:", ev["selected"], "| probe_error:", ev.get("probe_error"))
    ok = [r for r in rows if "ms_per_launch" in r]
    base = next((r["ms_per_launch"] for r in ok if r["name"] == "row_major"), None)
    for r in sorted(ok, key=lambda r: r["ms_per_launch"]):
        rel = "" if not base else "  x%.3f" % (base / r["ms_per_launch"])
        print("%9.3f ms%s  %s  %s" % (r["ms_per_launch"], rel, "identical" if r.get("identical_to_row_major") else "DIFFERENT", r["name"]))
    for r in rows:
        if "error" in r:
            print("    error  %s: %s" % (r["name"], r["error"]))


if __name__ == "__main__":
    main()
