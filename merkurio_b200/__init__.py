"""merkurio_b200 — B200-native k-mer / multi-pattern matching engine behind MerKurio's
`extract` / `tag` hot path. The product is the CUDA library (csrc/, C ABI in include/) and the
C++ host (host/); the Python here is plumbing for tests and benchmarks."""
__version__ = "0.2.0"
