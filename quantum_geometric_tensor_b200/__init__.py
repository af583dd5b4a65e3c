"""B200-native statevector + quantum-geometric-tensor hot path."""
__version__ = "0.1.0"
