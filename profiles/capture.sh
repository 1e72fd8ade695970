#!/bin/bash
# ncu captures for profiles/ (run under gpurun, 1 GPU). $1 = round tag, $2 = pop
TAG=${1:-r01}; POP=${2:-100000}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python profiles/run_cycles.py decks/c5g7/c5g7_2d $POP > gpurun_out/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_histories -s 4 -c 1 -f -o gpurun_out/prof_hist_$TAG \
    python profiles/run_cycles.py decks/c5g7/c5g7_2d $POP > gpurun_out/prof_hist_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lone -s 4 -c 1 -f -o gpurun_out/prof_lone_$TAG \
    python profiles/run_cycles.py decks/c5g7/c5g7_2d $POP > gpurun_out/prof_lone_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ce_lookup -s 2 -c 1 -f -o gpurun_out/prof_ce_$TAG \
    python profiles/run_ce_lookup.py 10000000 > gpurun_out/prof_ce_$TAG.log 2>&1
tail -n 2 gpurun_out/prof_hist_$TAG.log; tail -n 2 gpurun_out/prof_ce_$TAG.log
