#!/bin/bash
# Per-launch device times of one bench step (ncu, cold-cache, serialised: compare SHARES).
# usage (under gpurun): bash tools/launch_list.sh <tag> [extra bench args]
TAG=${1:-r01}; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/launches_${TAG}.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_${TAG}.csv > gpurun_out/launches_${TAG}_summary.txt
cat gpurun_out/launches_${TAG}_summary.txt | head -40
