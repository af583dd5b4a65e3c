"""One QGT evaluation of a named workload with device statistics and size-independent sanity properties
(Hermitian Q, positive semi-definite metric, antisymmetric Berry curvature).  For sizes too slow for bench.py's
warm-up rule, e.g. the 30-qubit / 256-parameter target:  python tools/run_once.py t30 [option=value ...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import numpy as np
from quantum_geometric_tensor_b200 import api, circuits as K

ctx = api.Context(0)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    ctx.set_option(k, float(v))
c = K.config(sys.argv[1])
th = K.default_angles(c.num_params)
t0 = time.perf_counter()
q = ctx.qgt(c, th)
wall = time.perf_counter() - t0
st = ctx.stats()
g, b = q.real, q.imag
w = np.linalg.eigvalsh(0.5 * (g + g.T))
out = {"workload": sys.argv[1], "qubits": c.num_qubits, "params": c.num_params, "wall_s": wall,
       "stats": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()},
       "hermitian_defect": float(np.abs(q - q.conj().T).max()), "metric_min_eig": float(w.min()), "metric_max_eig": float(w.max()),
       "metric_trace": float(np.trace(g)), "berry_antisymmetry_defect": float(np.abs(b + b.T).max())}
print(json.dumps(out))
