#!/bin/bash
# Round-end measurement on a GPU box (run from the repo root, e.g. `gpurun -- bash tools/profile.sh`):
# parity tests, the bench line of both arms, the ncu launch list of the bench command and one full
# capture of the dominant kernel, exported to CSV on the box (the .ncu-rep itself is too large to
# bring back).  Everything lands in gpurun_out/; the summaries judged are copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -3 gpurun_out/gpu_tests.log
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:search_persistent -s 2 -c 1 -o /tmp/prof \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2>/dev/null
ncu -i /tmp/prof.ncu-rep --page source --csv > gpurun_out/prof_source.csv 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])"
python -c "import json;d=json.load(open('gpurun_out/bench_reference.json'));print(d['value'], d['cpu_baseline'])"
ls -la gpurun_out
