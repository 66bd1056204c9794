DCB200_LIB=$PWD/clustering_b200/libdcb200_prof.so DCB200_GEMM_PROF=1 timeout 300 python scripts/profile_kernels.py C5 500000 1 2>&1 | grep -E "gscan|total|wait|pops_ms" | cut -c1-150
