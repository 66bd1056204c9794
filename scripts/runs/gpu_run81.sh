for w in C3 C2 C1; do for m in 17 3; do echo "== $w min_d $m"; DCB200_GEMM_MIN_D=$m timeout 300 python scripts/profile_kernels.py $w 2>&1 | tail -n 1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('pops', round(j['pops_ms'],2), 'nn', round(j['nn_ms'],2), 'nn_eval', round(j['nn_eval_frac'],3), 'nn_exact', j['nn_exact'], 'pops_exact', j['pops_exact'], 'max_pop', j['max_pop'])"; done; done
