"""Scan times of the FFMA kernels (DCB200_GEMM=0) and of the tensor-core path for a given n_cols on synthetic mixtures."""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clustering_b200.session import Session
from clustering_b200.synth import gaussian_mixture

n = int(sys.argv[1]); dims = [int(v) for v in sys.argv[2:]]
for d in dims:
    x = gaussian_mixture(n, d, k=12, seed=d)
    dd = ((x[:400, None, :] - x[None, :400, :]) ** 2).sum(-1)
    r = float(np.sqrt(np.percentile(dd[dd > 0], 3)))
    xd = torch.from_numpy(x).cuda()
    out = {"n": n, "d": d, "r": round(r, 3)}
    for mode in ("0", "1"):
        os.environ["DCB200_GEMM"] = mode
        s = Session(0)
        stream = s.torch_stream()
        for rep in range(2):
            s.set_coords(xd)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(stream)
            pp = s.populations([r])
            e[1].record(stream)
            fe = s.free_energies(s.to_frame_order(pp)[0].contiguous())
            s.nn_prepare(fe)
            e2 = torch.cuda.Event(enable_timing=True); e2.record(stream)
            keys = s.nn_scan()
            e[2].record(stream)
            e[2].synchronize()
        tag = "tensor" if s.gemm_info()[0] else "ffma"
        out[tag + "_pops_ms"] = round(e[0].elapsed_time(e[1]), 2)
        out[tag + "_nn_ms"] = round(e2.elapsed_time(e[2]), 2)
        out[tag + "_pops_sum"] = int(pp.sum().item())
        out[tag + "_keys"] = int(keys.sum().item() & 0xffffffff)
        s.close()
    print(json.dumps(out), flush=True)
