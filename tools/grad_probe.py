"""Device-time breakdown of the adjoint energy gradient against the forward circuit (tools, not a test).
usage: python tools/grad_probe.py [workload] [repeats]     QGT_B200_TRACE=1 prints every launch."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_geometric_tensor_b200 import api, circuits as K  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = K.config(name)
if not c.edges:
    n = c.num_qubits
    c.edges = [(q, (q + 1) % n, 1.0) for q in range(n)]
th = K.default_angles(c.num_params)
ctx = api.Context(0)
s = ctx.state(c.num_qubits).init(0)
for _ in range(reps):
    s.apply(c, th)
fwd = ctx.stats()
s.close()
for _ in range(reps):
    t0 = time.perf_counter()
    e, g = ctx.expectation_gradient(c, th)
    wall = (time.perf_counter() - t0) * 1e3
st = ctx.stats()
keys = ("ms_total", "ms_sweep", "ms_gram", "ms_other", "sweep_launches", "fused_launches", "other_launches", "num_runs", "fused")
print(name, "forward:", {k: round(fwd[k], 4) if isinstance(fwd[k], float) else fwd[k] for k in keys})
print(name, "gradient:", {k: round(st[k], 4) if isinstance(st[k], float) else st[k] for k in keys}, "wall ms", round(wall, 3))
print("ratio gradient / forward device time:", round(st["ms_total"] / fwd["ms_total"], 2), " E =", e, " |g| =", float(np.linalg.norm(g)))
