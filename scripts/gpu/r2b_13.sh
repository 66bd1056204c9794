#!/bin/bash
# round 2, session 2, call 13: neighbour-search knobs (seed width, window of the first pass) re-swept with the round-2 kernel
mkdir -p gpurun_out
timeout 400 python scripts/sweep_knobs.py C3 DCB200_NN_SEED_W=4 DCB200_NN_SEED_W=16 DCB200_NN_WINDOW=8 DCB200_NN_WINDOW=32 DCB200_NN_WINDOW=64 DCB200_NN_WINDOW=0 2>&1 | grep "^{" | cut -c1-170 | tee -a gpurun_out/r2b_13_sweep.jsonl
timeout 400 python scripts/sweep_knobs.py C2 DCB200_NN_SEED_W=4 DCB200_NN_SEED_W=16 DCB200_NN_WINDOW=8 DCB200_NN_WINDOW=32 DCB200_NN_WINDOW=64 2>&1 | grep "^{" | cut -c1-170 | tee -a gpurun_out/r2b_13_sweep.jsonl
