"""hairmsnn_b200 — B200-native (sm_100a) per-path rendering loop of HairMSNN.

The product is the C-ABI library `lib/libhairmsnn.so` (include/hairmsnn.h) and the three
headless executables under `bin/`; this package is the thin Python host used by the
tests and bench.py.  Import `hairmsnn_b200.api` to bind the library (raises if it has
not been built — there is no CPU fallback)."""
__all__ = ["api", "synth"]
