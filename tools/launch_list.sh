#!/bin/bash
# ncu launch list (device time of every kernel) of one bench.py run; output: gpurun_out/launches_<tag>.csv
tag=$1; shift
python bench.py "$@" > gpurun_out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$tag.csv python bench.py "$@" > gpurun_out/ncu_$tag.log 2>&1
