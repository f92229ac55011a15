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
# the probe of bench.py runs in a child process: force the evaluator for the captures (EVAL=row|column,
# KAO_SCHEDULE="sync,compress,threads,unroll" selects a schedule of the column-major one)
for EVAL in column row; do
timeout 500 ncu --set full --clock-control none --import-source on -k regex:search_persistent -s 2 -c 1 -o /tmp/prof_$EVAL \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --evaluator $EVAL > gpurun_out/ncu_full_$EVAL.log 2>&1
ncu -i /tmp/prof_$EVAL.ncu-rep --page raw --csv > gpurun_out/prof_raw_$EVAL.csv 2>/dev/null
ncu -i /tmp/prof_$EVAL.ncu-rep --page source --csv > gpurun_out/prof_source_$EVAL.csv 2>/dev/null
done
python -c "import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])"
python -c "import json;d=json.load(open('gpurun_out/bench_reference.json'));print(d['value'], d['cpu_baseline'])"
ls -la gpurun_out
