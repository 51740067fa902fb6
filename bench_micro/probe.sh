#!/bin/bash
# environment probe + integer peak microbenchmark (run under gpurun)
mkdir -p gpurun_out
{
  echo "== nproc"; nproc
  echo "== lscpu"; lscpu | head -25
  echo "== mem"; free -g | head -3
  echo "== nvidia-smi"; nvidia-smi
  nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,power.limit --format=csv
} > gpurun_out/probe_env.txt 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/int_peak_clocks.csv &
SMI=$!
./bench_micro/int_peak > gpurun_out/int_peak.jsonl 2> gpurun_out/int_peak.err
echo "int_peak exit $?"
kill $SMI
tail -40 gpurun_out/int_peak.jsonl
