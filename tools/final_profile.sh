#!/bin/bash
# Round-end evidence on one B200 (run through gpurun from the repo root):
#   tests, smoke, bench (own arm + reference arm), ncu launch list of the bench, one full ncu capture of a config-2
#   launch and one of each large-grid kernel (variants 5, 7 and 9).
# Outputs land in gpurun_out/ with the tag given as $1; tools/ncu_summary.py turns them into profiles/*.txt.
set -u
TAG=${1:-r2}
OUT=gpurun_out
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; tail -n 3 $OUT/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -n 1 $OUT/smoke_$TAG.log
python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "reference arm rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1; echo "ncu launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:hadi_douglas_kernel -s 3 -c 1 -f -o $OUT/prof_$TAG \
    python tools/prof_case.py 500 50 100 50 1 4 5 > $OUT/prof_$TAG.log 2>&1; echo "ncu full rc=$?"
HADI_FORCE_VARIANT=5 ncu --set full --clock-control none --import-source on -k regex:hadi_douglas_kernel -s 1 -c 1 -f -o $OUT/prof_v5_$TAG \
    python tools/prof_large.py 148 20 > $OUT/prof_v5_$TAG.log 2>&1; echo "ncu variant 5 rc=$?"
HADI_FORCE_VARIANT=7 ncu --set full --clock-control none --import-source on -k regex:hadi_cluster_kernel -s 1 -c 1 -f -o $OUT/prof_v7_$TAG \
    python tools/prof_large.py 1 20 > $OUT/prof_v7_$TAG.log 2>&1; echo "ncu variant 7 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:hadi_wide_kernel -s 1 -c 1 -f -o $OUT/prof_v9_$TAG \
    python tools/prof_large.py 1 20 > $OUT/prof_v9_$TAG.log 2>&1; echo "ncu variant 9 (wide kernel, one solve) rc=$?"
# text summaries on the box (gpurun copies back at most 64 MiB: the reports of variants 5 and 7 are summarised and dropped)
python tools/ncu_summary.py launches $OUT/launches_$TAG.csv > $OUT/sum_launches_$TAG.txt 2>&1
for v in "" _v5 _v7 _v9; do
  python tools/ncu_summary.py full $OUT/prof${v}_$TAG.ncu-rep > $OUT/sum_full${v}_$TAG.txt 2>&1
  python tools/ncu_lines.py $OUT/prof${v}_$TAG.ncu-rep 60 > $OUT/sum_lines${v}_$TAG.txt 2>&1
done
python tools/ncu_fp64_by_phase.py $OUT/prof_$TAG.ncu-rep 500 50 5151 72 > $OUT/sum_fp64_by_phase_$TAG.txt 2>&1
rm -f $OUT/prof_v5_$TAG.ncu-rep $OUT/prof_v7_$TAG.ncu-rep
ls -la $OUT/
