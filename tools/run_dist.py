"""One sharded QGT evaluation of a named workload, one process per GPU (torchrun), with device statistics and
size-independent sanity properties.  python -m torch.distributed.run --nproc-per-node 8 ... tools/run_dist.py c5 [opt=value ...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import numpy as np
import torch
import torch.distributed as dist
from quantum_geometric_tensor_b200 import api, circuits as K

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = api.Context(local)
uid = [api.Context.dist_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.dist_init(rank, world, uid[0])
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    ctx.set_option(k, float(v))
c = K.config(sys.argv[1])
th = K.default_angles(c.num_params)
dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
q = ctx.qgt(c, th)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
st = ctx.stats()
t = torch.tensor([st["ms_total"], st["ms_sweep"], st["ms_gram"], st["ms_exchange"], st["ms_other"]], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    g, b = q.real, q.imag
    w = np.linalg.eigvalsh(0.5 * (g + g.T))
    n = c.num_qubits
    out = {"workload": sys.argv[1], "qubits": n, "params": c.num_params, "gpus": world, "wall_s": wall,
           "ms_max_over_ranks": {"total": float(t[0]), "sweep": float(t[1]), "gram": float(t[2]), "exchange": float(t[3]), "other": float(t[4])},
           "stats_rank0": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()},
           "exchange_gbs_per_direction": st["exchange_bytes"] / max(st["ms_exchange"], 1e-9) * 1e-6,
           "hermitian_defect": float(np.abs(q - q.conj().T).max()), "metric_min_eig": float(w.min()), "metric_max_eig": float(w.max()),
           "metric_trace": float(np.trace(g)), "berry_antisymmetry_defect": float(np.abs(b + b.T).max()),
           "first_layer_ry_diag_defect": float(np.abs(np.diag(g)[:n] - 0.25).max()) if sys.argv[1] in ("c2", "c3", "c3s", "c5", "t30") else None,
           "first_layer_rz_diag_defect": float(np.abs(np.diag(g)[n:2 * n] - 0.25 * np.sin(th[:n]) ** 2).max()) if sys.argv[1] in ("c2", "c3", "c3s", "c5", "t30") else None}
    print(json.dumps(out), flush=True)
ctx.close()
dist.destroy_process_group()
