import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from quantum_geometric_tensor_b200 import api, circuits as K
ctx = api.Context(0)
for kv in sys.argv[2:]:
    k, v = kv.split("="); ctx.set_option(k, float(v))
c = K.config(sys.argv[1]); th = K.default_angles(c.num_params)
for i in range(3):
    if i == 2: sys.stderr.write("=== traced eval ===\n")
    q = ctx.qgt(c, th)
st = ctx.stats()
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
