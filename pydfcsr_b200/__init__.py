"""pydfcsr_b200 — B200-native CSR-wake hot path behind pyDFCSR's own call boundary."""
__version__ = "0.1.0"
