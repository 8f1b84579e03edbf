"""Host mirror of plonky2's `PolynomialBatch` / `MerkleTree` for the B200 backend
(upstream fri/oracle.rs, hash/merkle_tree.rs — the dependency pinned at
/root/reference/Cargo.toml:12).  Same names and argument meaning as the reference types; all
compute happens in libp2gpu.so."""
import ctypes as C
import numpy as np

from .ffi import load_library, P2GError


class Context:
    """One CUDA device context (p2g_ctx): a stream, NTT tables and a memory pool."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.p2g_ctx_create(device, C.byref(h))
        if rc != 0:
            raise P2GError(rc, "p2g_ctx_create: no usable CUDA device (there is no CPU fallback)")
        self.handle = h
        self.device = device

    def check(self, rc):
        if rc != 0:
            raise P2GError(rc, (self.lib.p2g_last_error(self.handle) or b"").decode())

    def sync(self):
        self.check(self.lib.p2g_ctx_sync(self.handle))

    @property
    def stream(self):
        return self.lib.p2g_ctx_stream(self.handle)

    def close(self):
        if self.handle:
            self.lib.p2g_ctx_destroy(self.handle)
            self.handle = None

    def poseidon_peak(self, iters=64):
        v = C.c_double()
        self.check(self.lib.p2g_poseidon_peak(self.handle, iters, C.byref(v)))
        return v.value

    def hash_no_pad_many(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out = np.empty((rows.shape[0], 4), dtype=np.uint64)
        self.check(self.lib.p2g_hash_no_pad_many(self.handle, rows.ctypes.data, rows.shape[0], rows.shape[1], out.ctypes.data))
        return out

    def merkle_cap(self, leaves, cap_height):
        """MerkleTree::new(leaves, cap_height).cap for row-major leaves; also returns leaf digests."""
        leaves = np.ascontiguousarray(leaves, dtype=np.uint64)
        n, ll = leaves.shape
        log_n = n.bit_length() - 1
        assert 1 << log_n == n
        cap = np.empty((1 << cap_height, 4), dtype=np.uint64)
        dig = np.empty((n, 4), dtype=np.uint64)
        self.check(self.lib.p2g_merkle_cap(self.handle, leaves.ctypes.data, log_n, ll, cap_height, cap.ctypes.data, dig.ctypes.data))
        return cap, dig


def _ptr(x):
    """host numpy array, or anything with .data_ptr() (a torch CUDA tensor), or a raw int"""
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return int(x)


class PolynomialBatch:
    """Device-resident PolynomialBatch: coefficients, LDE values (column-major, bit-reversed) and
    the Poseidon Merkle tree.  `from_values` / `from_coeffs` mirror the reference constructors
    (blinding is always false: zero_knowledge=false is the only supported configuration)."""

    def __init__(self, ctx, handle, ncols, log_n, rate_bits, cap_height, cap):
        self.ctx, self.handle = ctx, handle
        self.ncols, self.log_n, self.rate_bits, self.cap_height = ncols, log_n, rate_bits, cap_height
        self.cap = cap

    @classmethod
    def _make(cls, ctx, fn, cols, ncols, log_n, rate_bits, cap_height):
        h = C.c_void_p()
        cap = np.empty((1 << cap_height, 4), dtype=np.uint64)
        ctx.check(fn(ctx.handle, _ptr(cols), ncols, log_n, rate_bits, cap_height, C.byref(h), cap.ctypes.data))
        return cls(ctx, h, ncols, log_n, rate_bits, cap_height, cap)

    @classmethod
    def from_values(cls, ctx, values, rate_bits=3, cap_height=4):
        values = np.ascontiguousarray(values, dtype=np.uint64)
        ncols, n = values.shape
        return cls._make(ctx, ctx.lib.p2g_commit_from_values, values, ncols, n.bit_length() - 1, rate_bits, cap_height)

    @classmethod
    def from_coeffs(cls, ctx, coeffs, rate_bits=3, cap_height=4):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64)
        ncols, n = coeffs.shape
        return cls._make(ctx, ctx.lib.p2g_commit_from_coeffs, coeffs, ncols, n.bit_length() - 1, rate_bits, cap_height)

    @classmethod
    def from_values_device(cls, ctx, dev_cols, ncols, log_n, rate_bits=3, cap_height=4):
        return cls._make(ctx, ctx.lib.p2g_commit_from_values_dev, dev_cols, ncols, log_n, rate_bits, cap_height)

    @classmethod
    def from_coeffs_device(cls, ctx, dev_cols, ncols, log_n, rate_bits=3, cap_height=4):
        return cls._make(ctx, ctx.lib.p2g_commit_from_coeffs_dev, dev_cols, ncols, log_n, rate_bits, cap_height)

    @property
    def n(self):
        return 1 << self.log_n

    @property
    def lde_size(self):
        return 1 << (self.log_n + self.rate_bits)

    @property
    def path_len(self):
        return self.log_n + self.rate_bits - self.cap_height

    def coeffs(self):
        out = np.empty((self.ncols, self.n), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.p2g_batch_get_coeffs(self.ctx.handle, self.handle, out.ctypes.data))
        return out

    def lde_values(self):
        """[ncols][N], index j = evaluation at 7*w_N^bitrev(j) (the Merkle leaf order)."""
        out = np.empty((self.ncols, self.lde_size), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.p2g_batch_get_lde(self.ctx.handle, self.handle, out.ctypes.data))
        return out

    def digests(self, level):
        out = np.empty((self.lde_size >> level, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.p2g_batch_get_level(self.ctx.handle, self.handle, level, out.ctypes.data))
        return out

    def get_and_prove(self, leaf_index):
        """(MerkleTree::get(i), MerkleTree::prove(i).siblings)"""
        row = np.empty(self.ncols, dtype=np.uint64)
        sib = np.empty((self.path_len, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.p2g_batch_open_leaf(self.ctx.handle, self.handle, leaf_index, row.ctypes.data, sib.ctypes.data))
        return row, sib

    def free(self):
        if self.handle:
            self.ctx.lib.p2g_batch_free(self.ctx.handle, self.handle)
            self.handle = None
