"""Generates tests/golden/*.npz from the UNMODIFIED reference code (oracle/_ref/libqgt_ref.so, built by
oracle/Makefile from /root/reference) — run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

The vectors pin the oracle restatement and the CUDA path on boxes where the reference tree is absent.
Each file holds the circuit (gate table), theta, the final state from sim_execute_circuit
(hardware/quantum_simulator.c:499), g from diffgeo_compute_fubini_study and F = -2 Im Q from
diffgeo_compute_berry_curvature (distributed/differential_geometry.c:2819,2864).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.oracle import Reference  # noqa: E402
from quantum_geometric_tensor_b200 import circuits as K  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def gate_table(c):
    return np.array([[k, t, ctl, p, a, s] for (k, t, ctl, p, a, s) in c.gates], dtype=np.float64)


def main():
    ref = Reference()
    cases = {
        "hea_n4_l2": K.hea_layers(4, 2),
        "hea_n6_l1": K.hea_layers(6, 1),
        "hea_n8_p20": K.hea(8, 20),
        "rand_n5_refkinds": K.random_circuit(5, 40, 2024, kinds=sorted(K.REFERENCE_KINDS)),
        "rand_n7_refkinds": K.random_circuit(7, 60, 2025, kinds=sorted(K.REFERENCE_KINDS)),
        "c1_hea_n12_l2": K.config("c1"),
    }
    for name, c in cases.items():
        th = K.default_angles(c.num_params)
        psi = ref.apply(c, th)
        g, f = ref.qgt(c, th)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), gates=gate_table(c), n=c.num_qubits, theta=th,
                            psi=psi if c.num_qubits <= 8 else psi[:256], psi_len=psi.size,
                            psi_checksum=np.array([np.sum(psi * np.arange(1, psi.size + 1))]),
                            metric=g, curvature=f)
        print(name, c.num_qubits, c.num_params, "ok")


if __name__ == "__main__":
    main()
