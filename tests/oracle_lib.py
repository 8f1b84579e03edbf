"""ctypes loader for oracle/liboracle.so — the CPU checker (test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
P = 0xFFFFFFFF00000001


class Merkle(C.Structure):
    _fields_ = [("num_leaves", C.c_size_t), ("leaf_len", C.c_size_t), ("cap_height", C.c_int),
                ("leaves", C.c_void_p), ("digests", C.c_void_p), ("cap", C.c_void_p)]


class Batch(C.Structure):
    _fields_ = [("ncols", C.c_int), ("log_n", C.c_int), ("rate_bits", C.c_int), ("cap_height", C.c_int),
                ("coeffs", C.c_void_p), ("leaves", C.c_void_p), ("tree", C.POINTER(Merkle))]


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        vp = C.c_void_p
        lib.orc_poseidon.argtypes = [vp]
        lib.orc_poseidon_naive.argtypes = [vp]
        lib.orc_hash_n_to_m_no_pad.argtypes = [vp, C.c_size_t, vp, C.c_size_t]
        lib.orc_two_to_one.argtypes = [vp, vp, vp]
        for f in (lib.orc_fft, lib.orc_ifft):
            f.argtypes = [vp, C.c_int]
        for f in (lib.orc_coset_fft, lib.orc_coset_ifft):
            f.argtypes = [vp, C.c_int, C.c_uint64]
        for f in (lib.orc_batch_from_values, lib.orc_batch_from_coeffs):
            f.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
            f.restype = C.POINTER(Batch)
        lib.orc_batch_free.argtypes = [C.POINTER(Batch)]
        lib.orc_merkle_new.argtypes = [vp, C.c_size_t, C.c_size_t, C.c_int]
        lib.orc_merkle_new.restype = C.POINTER(Merkle)
        lib.orc_merkle_free.argtypes = [C.POINTER(Merkle)]
        lib.orc_merkle_level.argtypes = [C.POINTER(Merkle), C.c_int]
        lib.orc_merkle_level.restype = vp
        lib.orc_merkle_prove.argtypes = [C.POINTER(Merkle), C.c_size_t, vp]
        lib.orc_merkle_verify.argtypes = [vp, C.c_size_t, C.c_size_t, vp, C.c_int, vp, C.c_size_t]
        lib.orc_merkle_verify.restype = C.c_int
        lib.orc_num_threads.restype = C.c_int
        lib.orc_set_num_threads.argtypes = [C.c_int]
        lib.orc_circuit_load.argtypes = [vp]
        lib.orc_circuit_load.restype = vp
        lib.orc_circuit_free.argtypes = [vp]
        lib.orc_circuit_cap.argtypes = [vp]
        lib.orc_circuit_cap.restype = vp
        lib.orc_proof_len.argtypes = [vp]
        lib.orc_proof_len.restype = C.c_size_t
        lib.orc_prove_debug.argtypes = [vp, vp, vp, vp, C.c_size_t, vp, vp, vp]
        lib.orc_prove_debug.restype = C.c_long
        lib.orc_verify.argtypes = [vp, vp, C.c_size_t]
        lib.orc_verify.restype = C.c_int

    # --- small helpers returning numpy ---
    def poseidon(self, state):
        s = np.array(state, dtype=np.uint64)
        self.lib.orc_poseidon(s.ctypes.data)
        return s

    def poseidon_naive(self, state):
        s = np.array(state, dtype=np.uint64)
        self.lib.orc_poseidon_naive(s.ctypes.data)
        return s

    def hash_n_to_m_no_pad(self, inputs, m):
        a = np.ascontiguousarray(inputs, dtype=np.uint64)
        out = np.empty(m, dtype=np.uint64)
        self.lib.orc_hash_n_to_m_no_pad(a.ctypes.data, a.size, out.ctypes.data, m)
        return out

    def hash_no_pad(self, inputs):
        return self.hash_n_to_m_no_pad(inputs, 4)

    def two_to_one(self, l, r):
        l = np.ascontiguousarray(l, dtype=np.uint64); r = np.ascontiguousarray(r, dtype=np.uint64)
        out = np.empty(4, dtype=np.uint64)
        self.lib.orc_two_to_one(l.ctypes.data, r.ctypes.data, out.ctypes.data)
        return out

    def fft(self, a, inverse=False):
        a = np.array(a, dtype=np.uint64)
        (self.lib.orc_ifft if inverse else self.lib.orc_fft)(a.ctypes.data, a.size.bit_length() - 1)
        return a

    def coset_fft(self, a, shift=7, inverse=False):
        a = np.array(a, dtype=np.uint64)
        (self.lib.orc_coset_ifft if inverse else self.lib.orc_coset_fft)(a.ctypes.data, a.size.bit_length() - 1, shift)
        return a

    def batch(self, cols, from_values, rate_bits=3, cap_height=4):
        cols = np.ascontiguousarray(cols, dtype=np.uint64)
        ncols, n = cols.shape
        fn = self.lib.orc_batch_from_values if from_values else self.lib.orc_batch_from_coeffs
        return OracleBatch(self, fn(cols.ctypes.data, ncols, n.bit_length() - 1, rate_bits, cap_height))

    def merkle(self, leaves, cap_height):
        leaves = np.ascontiguousarray(leaves, dtype=np.uint64)
        t = self.lib.orc_merkle_new(leaves.ctypes.data, leaves.shape[0], leaves.shape[1], cap_height)
        return OracleMerkle(self, t, leaves)


def _view(ptr, shape):
    n = int(np.prod(shape))
    buf = (C.c_uint64 * n).from_address(ptr)
    return np.frombuffer(buf, dtype=np.uint64).reshape(shape).copy()


class OracleMerkle:
    def __init__(self, orc, t, leaves=None, own=True):
        self.orc, self.t, self.leaves, self.own = orc, t, leaves, own
        self.num_leaves = t.contents.num_leaves
        self.cap_height = t.contents.cap_height
        self.path_len = (self.num_leaves.bit_length() - 1) - self.cap_height

    def level(self, k):
        cnt = self.num_leaves >> k
        return _view(self.orc.lib.orc_merkle_level(self.t, k), (cnt, 4))

    @property
    def cap(self):
        return self.level(self.path_len)

    def prove(self, i):
        sib = np.empty((max(self.path_len, 1), 4), dtype=np.uint64)
        self.orc.lib.orc_merkle_prove(self.t, i, sib.ctypes.data)
        return sib[:self.path_len]

    def free(self):
        if self.own and self.t:
            self.orc.lib.orc_merkle_free(self.t)
            self.t = None


class OracleBatch:
    def __init__(self, orc, b):
        self.orc, self.b = orc, b
        c = b.contents
        self.ncols, self.log_n, self.rate_bits, self.cap_height = c.ncols, c.log_n, c.rate_bits, c.cap_height
        self.n = 1 << c.log_n
        self.N = self.n << c.rate_bits
        self.tree = OracleMerkle(orc, c.tree, own=False)

    def coeffs(self):
        return _view(self.b.contents.coeffs, (self.ncols, self.n))

    def leaves(self):
        """row-major [N][ncols] as upstream stores them"""
        return _view(self.b.contents.leaves, (self.N, self.ncols))

    def free(self):
        if self.b:
            self.orc.lib.orc_batch_free(self.b)
            self.b = None


class OracleTranscript(C.Structure):
    _fields_ = [("betas", C.c_uint64 * 4), ("gammas", C.c_uint64 * 4), ("deltas", C.c_uint64 * 16),
                ("alphas", C.c_uint64 * 4), ("zeta", C.c_uint64 * 2), ("fri_alpha", C.c_uint64 * 2),
                ("fri_betas", C.c_uint64 * 32), ("pow_witness", C.c_uint64), ("query_indices", C.c_uint64 * 64)]


def set_circuit_digest(orc, data):
    """circuit_digest = hash_no_pad(constants_sigmas cap || hash_pad([]) || degree_bits) computed with the
    oracle (CircuitData.load does the same on the device); needed before proving on the CPU only."""
    oc = OracleCircuit(orc, data)
    cap = oc.cap.copy()
    oc.free()
    lib = orc.lib
    lib.orc_hash_no_pad.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    pad = np.array([1] + [0] * 10 + [1], dtype=np.uint64)
    dom = np.zeros(4, dtype=np.uint64)
    lib.orc_hash_no_pad(pad.ctypes.data, pad.size, dom.ctypes.data)
    parts = np.concatenate([cap.ravel(), dom, np.array([data.degree_bits], dtype=np.uint64)])
    dig = np.zeros(4, dtype=np.uint64)
    lib.orc_hash_no_pad(parts.ctypes.data, parts.size, dig.ctypes.data)
    data.circuit_digest = dig
    data.constants_sigmas_cap = cap
    return dig


class OracleCircuit:
    """orc_circuit_load on the descriptor of a host CircuitData (same C layout as p2g_circuit_desc)."""

    def __init__(self, orc, data):
        self.orc, self.data = orc, data
        self.desc = data.descriptor()
        self.h = orc.lib.orc_circuit_load(C.addressof(self.desc))
        self.proof_len = orc.lib.orc_proof_len(C.addressof(self.desc))
        self.cap = _view(orc.lib.orc_circuit_cap(self.h), (1 << data.config.fri_config.cap_height, 4))

    def prove(self, wires, debug=False, public_inputs=None):
        wires = np.ascontiguousarray(wires, dtype=np.uint64)
        pi = np.ascontiguousarray(public_inputs, dtype=np.uint64) if self.desc.num_public_inputs else None
        assert pi is None or pi.size == self.desc.num_public_inputs
        out = np.zeros(self.proof_len, dtype=np.uint64)
        tr = OracleTranscript()
        zs = qc = None
        zs_p = qc_p = None
        if debug:
            d = self.desc
            nlp = 0 if d.num_luts == 0 else -(-(d.num_routed_wires // 2) // (d.quotient_degree_factor - 1)) + 1
            zs_cols = d.num_challenges * (1 + d.num_partial_products + nlp)
            zs = np.zeros((zs_cols, self.data.n), dtype=np.uint64)
            qc = np.zeros((d.num_challenges * d.quotient_degree_factor, self.data.n), dtype=np.uint64)
            zs_p, qc_p = zs.ctypes.data, qc.ctypes.data
        rc = self.orc.lib.orc_prove_debug(self.h, wires.ctypes.data, pi.ctypes.data if pi is not None else None, out.ctypes.data, out.size, C.addressof(tr), zs_p, qc_p)
        if rc < 0:
            raise ValueError(f"oracle prove failed: {rc}")
        assert rc == self.proof_len, (rc, self.proof_len)
        return (out, tr, zs, qc) if debug else out

    def verify(self, proof):
        proof = np.ascontiguousarray(proof, dtype=np.uint64)
        return self.orc.lib.orc_verify(self.h, proof.ctypes.data, proof.size)

    def free(self):
        if self.h:
            self.orc.lib.orc_circuit_free(self.h)
            self.h = None


_ORACLE = None


def load():
    global _ORACLE
    if _ORACLE is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h", ".inc"))]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
        _ORACLE = Oracle(C.CDLL(so))
    return _ORACLE
