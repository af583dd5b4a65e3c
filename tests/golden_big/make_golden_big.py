"""Generates tests/golden_big/*.npz: SUB-BLOCKS of the tensor at the sizes the blocked / fused schedules actually run
at, from the UNMODIFIED reference code where it has the gates (oracle/_ref: sim_execute_circuit for every column,
diffgeo_compute_fubini_study / _berry_curvature for the assembly) and from the oracle restatement for the QAOA
instance (the reference simulator has no cost-layer gate).  Run in the build container (several minutes of CPU):

    python tests/golden_big/make_golden_big.py

Kept apart from tests/golden/ because the loader there globs every file as a full-tensor case.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.oracle import Oracle, Reference  # noqa: E402
from quantum_geometric_tensor_b200 import circuits as K  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sub_block(eng, c, th, cols, reference):
    psi = eng.apply(c, th)
    J = np.stack([eng.derivative(c, th, mu) for mu in cols])
    if reference:
        g, f = eng.fubini_berry(psi, J)              # g = Re Q, f = -2 Im Q exactly as the reference returns them
    else:
        q = eng.qgt_from_columns(psi, J)
        g, f = q.real, -2 * q.imag
    v = np.array([np.vdot(J[i], psi) for i in range(len(cols))])          # <d_mu psi|psi>, for diagnostics
    return g, f, v, float(np.vdot(psi, psi).real)


def main():
    ref, orc = Reference(), Oracle()
    cases = [
        # C3's parameter count on 24 qubits: the column-block and the fused schedules are forced at this size in the tests
        ("c3s_hea_n24_p256_cols", K.config("c3s"), [0, 100, 255], ref, True),
        # 22 qubits, 3 full layers: first / middle / last layer, for the sharded multi-GPU check (>= 2 exchanges per rank qubit)
        ("hea_n22_l3_cols", K.hea_layers(22, 3), [5, 60, 131], ref, True),
        # QAOA MaxCut on 20 qubits, p = 3 (cost-layer generator + shared mixer parameter): oracle restatement
        ("qaoa_n20_p3_cols", K.qaoa_maxcut(20, 3), [0, 3, 5], orc, False),
    ]
    for name, c, cols, eng, is_ref in cases:
        t0 = time.time()
        th = K.default_angles(c.num_params)
        g, f, v, nrm = sub_block(eng, c, th, cols, is_ref)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), workload=name, n=c.num_qubits, num_params=c.num_params, cols=np.array(cols),
                            theta=th, metric=g, curvature=f, proj=v, norm2=nrm, source="oracle/_ref (unmodified reference)" if is_ref else "oracle port")
        print(name, c.num_qubits, c.num_params, cols, "%.0f s" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
