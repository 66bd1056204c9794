set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r70_bench_c2_2gpu.json 2> gpurun_out/r70_c2.err; tail -n 3 gpurun_out/r70_c2.err; cat gpurun_out/r70_bench_c2_2gpu.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 --workload C5 > gpurun_out/r70_bench_c5_2gpu.json 2> gpurun_out/r70_c5.err; tail -n 3 gpurun_out/r70_c5.err; cat gpurun_out/r70_bench_c5_2gpu.json | cut -c1-300
timeout 300 python - <<'PY'
import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
# single-process check that DensityPass (world 1) still equals the host-pointer API
from clustering_b200 import density
from clustering_b200.session import Session
from clustering_b200.dist import DensityPass
from clustering_b200.synth import gaussian_mixture
x = gaussian_mixture(50000, 5, seed=3)
s = Session(0)
with torch.cuda.stream(s.torch_stream()):
    pops, fe, nn = DensityPass(s, len(x), np.array([0.3], np.float32)).run(x)
s.sync()
p2 = density.calculate_populations(x, [0.3])
assert np.array_equal(pops.cpu().numpy().astype(np.uint32), p2)
n2 = density.nearest_neighbors(x, density.calculate_free_energies(p2[0]))
assert all(np.array_equal(a.cpu().numpy().view(np.uint32), b.view(np.uint32)) for a, b in zip(nn, n2))
print("DensityPass world=1 ok")
PY
