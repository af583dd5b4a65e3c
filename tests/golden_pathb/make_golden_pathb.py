"""Generates tests/golden_pathb/*.npz from the UNMODIFIED reference (oracle/_ref/libqgt_refb.so =
core/quantum_circuit_operations.c compiled where it lies by oracle/Makefile) — run in the build container:

    python tests/golden_pathb/make_golden_pathb.py

Each file: the circuit table (op, qubit a, qubit b, angle; op codes in tests/pathb.py), the qubit count and the final
ComplexFloat amplitudes of quantum_circuit_execute (quantum_circuit_operations.c:1139-1197) on |0...0>."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pathb  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = pathb.PathB(pathb.REFB)
    cases = {"rand_n1": (1, 12, 11), "rand_n3": (3, 30, 12), "rand_n6": (6, 80, 13), "rand_n10": (10, 150, 14), "rand_n14": (14, 200, 15)}
    for name, (n, ng, seed) in cases.items():
        table = pathb.random_table(n, ng, seed)
        rc, psi, stats = ref.run(n, table)
        assert rc == 0
        np.savez_compressed(os.path.join(HERE, name + ".npz"), n=n, table=table, psi=psi, depth=stats[0], gate_count=stats[1])
        print(name, n, ng, "depth", stats[0], "norm", float(np.vdot(psi, psi).real))
    # every named gate alone on a superposed register: pins the phase conventions (X = RX(pi) etc.)
    n = 3
    prep = [(pathb.H, q, 0, 0.0) for q in range(n)] + [(pathb.RZ, 1, 0, 0.7), (pathb.RY, 2, 0, -0.4)]
    for op in range(11):
        row = (op, 1, 2 if op >= pathb.CNOT else 0, 0.9 if op in (pathb.RX, pathb.RY, pathb.RZ) else (np.pi / 2 if op == pathb.PHASE else 0.0))
        table = np.array(prep + [row], dtype=np.float64)
        rc, psi, stats = ref.run(n, table)
        assert rc == 0
        np.savez_compressed(os.path.join(HERE, "single_" + pathb.NAMES[op] + ".npz"), n=n, table=table, psi=psi, depth=stats[0], gate_count=stats[1])
    print("single-gate cases ok")


if __name__ == "__main__":
    main()
