#!/bin/bash
# round 2, first light of the table-driven multi-radius population kernel: parity, then C3 / C1 timings per mode
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "bin_mode or kernels_agree or c3_shape" 2>&1 | tail -15
for mode in hist bin; do
  echo "== C3 mode $mode"; DCB200_POPS_MODE=$mode timeout 300 python scripts/profile_kernels.py C3 1000000 2 2>&1 | tail -n 1
done
for dl in 1 4 16 33; do
  echo "== C3 bin dense_lanes $dl"; DCB200_BIN_DENSE_LANES=$dl timeout 300 python scripts/profile_kernels.py C3 1000000 2 2>&1 | tail -n 1
done
for mode in count bin; do
  echo "== C1 mode $mode"; DCB200_POPS_MODE=$mode timeout 300 python scripts/profile_kernels.py C1 100000 3 2>&1 | tail -n 1
done
