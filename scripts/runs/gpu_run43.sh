for m in 0 3; do echo "== debug mode $m"; DCB200_GEMM_DEBUG=$m DCB200_GEMM_PROF=1 timeout 300 python scripts/profile_kernels.py C5 200000 2 2>&1 | tail -n 14 | grep -E "total|wait|pops_ms" | cut -c1-200; done
DCB200_GEMM_RA=1 DCB200_GEMM_PROF=1 timeout 300 python scripts/profile_kernels.py C5 200000 2 2>&1 | tail -n 14 | grep -E "total|wait|pops_ms" | cut -c1-200
timeout 600 python scripts/profile_kernels.py C5 500000 2 2>&1 | tail -n 1 | cut -c1-400
