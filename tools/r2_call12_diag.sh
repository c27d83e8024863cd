#!/bin/bash
# Round 2, GPU call 12 (1 GPU, < 2 min): why did bench.py --memory-type distributed end without a line for the exchange path and for
# the reference in call 11?  Same command at world size 1 (self-exchange) with stderr kept, bounded by short timeouts.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
A="--gpus 1 --memory-type distributed --no-e2e --no-cpu-baseline --steps 3 --warmup 3 --rows-per-gpu 2000000 --dim 256 --dtype fp32"
WG_FORCE_EXCHANGE=1 timeout 50 python bench.py $A > gpurun_out/diag_exchange.out 2> gpurun_out/diag_exchange.err; echo "exchange rc=$?"; tail -c 300 gpurun_out/diag_exchange.out; tail -6 gpurun_out/diag_exchange.err | cut -c1-300
timeout 50 python bench.py $A --impl reference > gpurun_out/diag_reference.out 2> gpurun_out/diag_reference.err; echo "reference rc=$?"; tail -c 300 gpurun_out/diag_reference.out; tail -6 gpurun_out/diag_reference.err | cut -c1-300
