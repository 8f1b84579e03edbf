"""plonky2_aes_b200 — B200-native backend for the plonky2 prove() hot path used by the
0xPARC/plonky2-aes gadget crates.  The package is a thin host layer over libp2gpu.so (CUDA,
sm_100a); there is no CPU fallback: anything that computes raises if the library or a GPU is
missing."""
from .host.ffi import lib_path, load_library, P2GError  # noqa: F401
