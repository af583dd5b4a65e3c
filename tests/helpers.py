"""Shared helpers for the test-suite."""
import glob
import os

import numpy as np

from quantum_geometric_tensor_b200 import circuits as K

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    c = K.Circuit(int(z["n"]))
    for row in z["gates"]:
        c.add(int(row[0]), int(row[1]), int(row[2]), int(row[3]), float(row[4]), float(row[5]))
    return c, z


def rel_err(a, b):
    d = np.abs(np.asarray(a) - np.asarray(b)).max()
    s = np.abs(np.asarray(b)).max()
    return d / s if s > 0 else d


# known-answer cases of the reference's tests/test_quantum_simulator_cpu.c (magnitudes, 1e-10)
def known_answer_cases():
    s = 0.70710678118654752440
    cases = []
    c = K.Circuit(1); c.add(K.H, 0)
    cases.append(("hadamard :64-67", c, [s, s]))
    c = K.Circuit(2); c.add(K.H, 0); c.add(K.CNOT, 1, 0)
    cases.append(("bell :131-134", c, [s, 0, 0, s]))
    c = K.Circuit(1); c.add(K.X, 0)
    cases.append(("pauli-x :177-178", c, [0, 1]))
    c = K.Circuit(1); c.add(K.H, 0); c.add(K.Z, 0)
    cases.append(("h-z :234-235", c, [s, s]))
    c = K.Circuit(1); c.add(K.RX, 0, -1, -1, np.pi)
    cases.append(("rx(pi) :281-282", c, [0, 1]))
    c = K.Circuit(3); c.add(K.H, 0); c.add(K.CNOT, 1, 0); c.add(K.CNOT, 2, 1)
    cases.append(("ghz :404-408", c, [s, 0, 0, 0, 0, 0, 0, s]))
    return cases
