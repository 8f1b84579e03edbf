"""Proof container: named views over the flat u64 proof words and the byte
serialisation of plonky2's `ProofWithPublicInputs::to_bytes` / `from_bytes`
(util/serialization/mod.rs of the pinned dependency; the byte format itself is implemented once, in the C ABI:
p2g_proof_to_bytes / p2g_proof_from_bytes, include/p2gpu.h).
The reference never calls the serialiser; it is provided for shipping
proofs off the box."""
import ctypes as C

import numpy as np

from .ffi import P2GError, load_library


def _layout(d):
    """yield (name, count_words, kind) segments in proof order; kind 'len' marks a Merkle-path length word"""
    cap = 4 << d.cap_height
    nch = d.num_challenges
    nlp = 0 if d.num_luts == 0 else -(-(d.num_routed_wires // 2) // (d.quotient_degree_factor - 1)) + 1
    NC = d.num_selectors + d.num_lookup_selectors + d.num_constants
    logN = d.degree_bits + d.rate_bits
    zs_cols = nch * (1 + d.num_partial_products + nlp)
    yield "wires_cap", cap, "f"
    yield "plonk_zs_partial_products_cap", cap, "f"
    yield "quotient_polys_cap", cap, "f"
    for name, cnt in (("constants", NC), ("plonk_sigmas", d.num_routed_wires), ("wires", d.num_wires), ("plonk_zs", nch),
                      ("plonk_zs_next", nch), ("partial_products", nch * d.num_partial_products),
                      ("quotient_polys", nch * d.quotient_degree_factor), ("lookup_zs", nch * nlp), ("lookup_zs_next", nch * nlp)):
        yield "openings." + name, 2 * cnt, "f"
    nl = d.num_reduction_arity_bits
    for l in range(nl):
        yield f"fri.commit_phase_merkle_caps[{l}]", cap, "f"
    cols = [NC + d.num_routed_wires, d.num_wires, zs_cols, nch * d.quotient_degree_factor]
    for q in range(d.num_query_rounds):
        for o in range(4):
            pl = logN - d.cap_height
            yield f"fri.query[{q}].initial[{o}].evals", cols[o], "f"
            yield f"fri.query[{q}].initial[{o}].path_len", 1, "len"
            yield f"fri.query[{q}].initial[{o}].siblings", 4 * pl, "f"
        lg = logN
        for l in range(nl):
            ab = d.reduction_arity_bits[l]
            lg -= ab
            yield f"fri.query[{q}].step[{l}].evals", 2 << ab, "f"
            yield f"fri.query[{q}].step[{l}].path_len", 1, "len"
            yield f"fri.query[{q}].step[{l}].siblings", 4 * (lg - d.cap_height), "f"
    fin = 1 << d.degree_bits
    for l in range(nl):
        fin >>= d.reduction_arity_bits[l]
    yield "fri.final_poly", 2 * fin, "f"
    yield "fri.pow_witness", 1, "f"
    yield "public_inputs", d.num_public_inputs, "f"


class Proof:
    """ProofWithPublicInputs over the flat words produced by p2g_prove / the oracle."""

    def __init__(self, words, desc):
        self.words = np.ascontiguousarray(words, dtype=np.uint64)
        self.desc = desc
        self.segments = {}
        pos = 0
        for name, cnt, kind in _layout(desc):
            self.segments[name] = (pos, cnt, kind)
            pos += cnt
        if pos != self.words.size:
            raise ValueError(f"proof has {self.words.size} words, layout expects {pos}")

    def __getitem__(self, name):
        pos, cnt, _ = self.segments[name]
        return self.words[pos:pos + cnt]

    @property
    def pow_witness(self):
        return int(self["fri.pow_witness"][0])

    def to_bytes(self):
        """ProofWithPublicInputs::to_bytes through the C ABI (p2g_proof_to_bytes; host-only code of libp2gpu.so)"""
        lib = load_library()
        n = lib.p2g_proof_bytes_len(C.byref(self.desc))
        out = np.empty(n, dtype=np.uint8)
        got = C.c_size_t()
        rc = lib.p2g_proof_to_bytes(C.byref(self.desc), self.words.ctypes.data, self.words.size, out.ctypes.data, n, C.byref(got))
        if rc != 0:
            raise P2GError(rc, "p2g_proof_to_bytes")
        return out[:got.value].tobytes()

    @classmethod
    def from_bytes(cls, data, desc):
        """ProofWithPublicInputs::from_bytes(bytes, common_data): raises on malformed input"""
        lib = load_library()
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        nwords = sum(cnt for _, cnt, _ in _layout(desc))
        words = np.empty(nwords, dtype=np.uint64)
        got = C.c_size_t()
        rc = lib.p2g_proof_from_bytes(C.byref(desc), buf.ctypes.data, buf.size, words.ctypes.data, nwords, C.byref(got))
        if rc != 0:
            raise ValueError("malformed proof bytes")
        return cls(words[:got.value], desc)
