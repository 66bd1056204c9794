#!/bin/bash
# knob sweeps of the round and the shards of an 8-GPU run timed on one GPU
mkdir -p gpurun_out
timeout 400 python scripts/sweep_knobs.py C3 DCB200_BIN_STEAL=0 DCB200_BIN_DENSE_LANES=4 DCB200_BIN_DENSE_LANES=16 DCB200_ITEMS_PER_CTA=48 DCB200_ITEMS_PER_CTA=384 DCB200_BIN_PROJ=2 DCB200_SUPER_PRUNE=0 DCB200_NN_SEED_W=4 DCB200_NN_SEED_W=16 DCB200_NN_WINDOW=4 DCB200_NN_WINDOW=16 DCB200_NN_WINDOW=32 2>&1 | grep "^{" | tee -a gpurun_out/sweep_c3.jsonl
timeout 300 python scripts/shard_timing.py C3 8 0 3 2>&1 | tail -n 1 | tee gpurun_out/shards8.json
DCB200_ITEMS_PER_CTA=1 timeout 300 python scripts/shard_timing.py C3 8 0 3 2>&1 | tail -n 1 | tee gpurun_out/shards8_16ranges.json
