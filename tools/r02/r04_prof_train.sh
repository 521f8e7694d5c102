#!/bin/bash
# r04: torch-profiler table of one config-5 training step (batch 12, bf16 autocast, channels-last net_g)
T=${1:-r04d}
mkdir -p gpurun_out
timeout 600 python tools/prof_train_step.py --bf16 --channels-last > gpurun_out/${T}_train_prof.txt 2>&1
tail -45 gpurun_out/${T}_train_prof.txt | cut -c1-220
