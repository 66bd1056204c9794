#!/bin/bash
# host-assembly path after the device fix; the command-line binary on 2 GPUs against 1 GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -12
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from clustering_b200.synth import config_data
x = config_data("C3", 200000)
np.savetxt("/tmp/c3s.coords", x, fmt="%.6f")
PY
R="0.1 0.2 0.3 0.4 0.5 0.6 0.7 0.8 0.9 1.0 1.1 1.2 1.3 1.4 1.5 1.6 1.7 1.8 1.9 2.0"
time env DCB200_TRACE=1 ./clustering_b200/clustering density -f /tmp/c3s.coords -R $R -p /tmp/pop -d /tmp/fe 2>&1 | tail -6
time env DCB200_TRACE=1 ./clustering_b200/clustering density -f /tmp/c3s.coords -r 1.0 -p /tmp/pop1 -d /tmp/fe1 -b /tmp/nn1 2>&1 | tail -6
CUDA_VISIBLE_DEVICES=0 ./clustering_b200/clustering density -f /tmp/c3s.coords -R $R -p /tmp/pop_1g -d /tmp/fe_1g 2>&1 | tail -2
CUDA_VISIBLE_DEVICES=0 ./clustering_b200/clustering density -f /tmp/c3s.coords -r 1.0 -p /tmp/pop1_1g -d /tmp/fe1_1g -b /tmp/nn1_1g 2>&1 | tail -2
ls /tmp | grep -E "^pop|^fe|^nn" | head -50 | tr '\n' ' '
# the headers (#@ lines) carry the command line, i.e. the differing output names: compare the data lines
same=1; for p in "pop_2.000000 pop_1g_2.000000" "fe_1.000000 fe_1g_1.000000" "nn1 nn1_1g" "pop1 pop1_1g"; do set -- $p; cmp <(grep -v '^#' /tmp/$1) <(grep -v '^#' /tmp/$2) || same=0; done
[ $same = 1 ] && echo "2-GPU files == 1-GPU files (data lines)"
