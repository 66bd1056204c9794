for it in 16 4 8 32 64; do echo "== column items $it"; DCB200_GEMM_ITEMS=$it timeout 300 python scripts/profile_kernels.py C5 500000 3 2>&1 | tail -n 2 | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('pops', round(j['pops_ms'],2), 'nn', round(j['nn_ms'],2), 'nn_eval', round(j['nn_eval_frac'],3), 'exact', j['nn_exact'])"; done
