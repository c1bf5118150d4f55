#!/bin/bash
# DRAM bytes and duration of every k_value_mma launch of one step, per workload shape (ncu, clocks untouched).
# Run on the GPU box from the repo root; writes gpurun_out/traffic_<name>.csv.  profiles/tools/traffic_to_json.py turns the
# CSVs into profiles/value_kernel_traffic.json (what bench.py's roofline.traffic reads).
set -e
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
run() { # name P N F
  python profiles/exp/prof_run.py P=$2 N=$3 F=$4 ${5:+alpha=$5} > /dev/null
  ncu --metrics $M --clock-control none -k regex:k_value_mma --csv --log-file gpurun_out/traffic_$1.csv python profiles/exp/prof_run.py P=$2 N=$3 F=$4 ${5:+alpha=$5} > gpurun_out/traffic_$1.log 2>&1
}
run c2 10000 100000 32
run c3_125000 125000 10000 64
run c3_500000 500000 10000 64
run c3_250000 250000 10000 64
run c3_1000000 1000000 10000 64
