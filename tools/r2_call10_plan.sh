#!/bin/bash
# Round 2, GPU call 10 (1 GPU): the shipped launch plan's neighbourhood on BENCH-SIZED tables (the first sweeps used 5 GB tables).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=wholegraph_b200/lib/rowmove_lab
timeout 300 $L --row-bytes 1024 --rows 100000000 --mode local --set plan | tee gpurun_out/lab_plan_rb1024.txt
timeout 300 $L --row-bytes 512 --rows 125000000 --mode local --set plan | tee gpurun_out/lab_plan_rb512.txt
timeout 300 $L --row-bytes 256 --rows 125000000 --mode local --set plan | tee gpurun_out/lab_plan_rb256.txt
