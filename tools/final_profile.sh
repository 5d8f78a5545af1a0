#!/bin/bash
# Round-end evidence on one B200 (run through gpurun from the repo root):
#   tests, bench (own arm + reference arm), ncu launch list of the bench, one full ncu capture of a config-2 launch.
# Outputs land in gpurun_out/ with the tag given as $1; tools/ncu_summary.py turns them into profiles/*.txt.
set -u
TAG=${1:-r1s4}
OUT=gpurun_out
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; tail -n 3 $OUT/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "reference arm rc=$?"
python tools/time_config4.py > $OUT/config4_$TAG.log 2>&1; echo "config4 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 > $OUT/ncu_bench_$TAG.log 2>&1; echo "ncu launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:hadi_douglas_kernel -s 3 -c 1 -f -o $OUT/prof_$TAG \
    python tools/prof_case.py 500 50 100 50 1 4 5 > $OUT/prof_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT/prof_$TAG.ncu-rep
