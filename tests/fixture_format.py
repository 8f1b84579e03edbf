"""Reference fixture files (`*.p2gfix`): reader, writer and the descriptor built from one.

Written by tools/dump_reference_fixture.rs on a machine that has the Rust reference
(plonky2 @ 109d517 through /root/reference/Cargo.toml:12); the file carries everything the hot path needs
-- circuit description as CircuitBuilder::build left it, the full wire matrix of one witness and the bytes
of the proof the REAL prover produced -- so the oracle and the CUDA path can be compared with the reference
bit for bit without Rust.  Format: magic "P2GFIX1\\0", u32 section count, then per section u32 name length,
name, u32 element size, u64 element count, raw little-endian data."""
import ctypes as C
import struct

import numpy as np

from plonky2_aes_b200.host import ffi

MAGIC = b"P2GFIX1\0"
_DT = {1: np.uint8, 2: np.uint16, 4: np.int32, 8: np.uint64}
CONFIG_FIELDS = ("degree_bits", "num_wires", "num_routed_wires", "num_constants", "num_challenges", "quotient_degree_factor",
                 "rate_bits", "cap_height", "pow_bits", "num_query_rounds", "num_selectors", "num_lookup_selectors",
                 "num_gate_constraints", "num_partial_products", "num_public_inputs")


def read(path):
    raw = open(path, "rb").read()
    if raw[:8] != MAGIC:
        raise ValueError(f"{path}: not a P2GFIX1 file")
    (count,), off = struct.unpack_from("<I", raw, 8), 12
    out = {}
    for _ in range(count):
        (nl,) = struct.unpack_from("<I", raw, off); off += 4
        name = raw[off:off + nl].decode(); off += nl
        esz, cnt = struct.unpack_from("<IQ", raw, off); off += 12
        out[name] = np.frombuffer(raw, dtype=np.dtype(_DT[esz]).newbyteorder("<"), count=cnt, offset=off).astype(_DT[esz])
        off += esz * cnt
    if off != len(raw):
        raise ValueError(f"{path}: trailing bytes")
    return out


def write(path, sections):
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<I", len(sections)))
        for name, arr in sections.items():
            arr = np.ascontiguousarray(arr)
            f.write(struct.pack("<I", len(name)) + name.encode() + struct.pack("<IQ", arr.dtype.itemsize, arr.size))
            f.write(arr.astype(arr.dtype.newbyteorder("<")).tobytes())


class FixtureCircuit:
    """The duck type tests/oracle_lib.OracleCircuit and the GPU loader expect (descriptor(), n, config ...),
    built from a fixture instead of this repo's circuit builder."""

    def __init__(self, fx):
        self.fx = fx
        self.cfg = dict(zip(CONFIG_FIELDS, (int(v) for v in fx["config"])))
        self.degree_bits = self.cfg["degree_bits"]
        self.n = 1 << self.degree_bits
        self.gates = fx["gates"].reshape(-1, 6)
        ids = fx["gate_ids"].tobytes().decode().split("\n")
        self.unsupported = [ids[i] if i < len(ids) else str(i) for i, g in enumerate(self.gates) if g[0] < 0]
        self._keep = {}

        class _Fri:
            cap_height = self.cfg["cap_height"]
        class _Cfg:
            fri_config = _Fri
        self.config = _Cfg
        self.circuit_digest = fx["circuit_digest"].astype(np.uint64)
        self.constants_sigmas_cap = fx["constants_sigmas_cap"].reshape(-1, 4).astype(np.uint64)
        self.wires = fx["wires"].reshape(self.cfg["num_wires"], self.n).astype(np.uint64)
        self.public_inputs = fx["public_inputs"].astype(np.uint64)
        self.proof_bytes = fx["proof_bytes"].tobytes()

    def descriptor(self):
        d, c, k = ffi.CircuitDesc(), self.cfg, self._keep
        for f in CONFIG_FIELDS:
            setattr(d, f, c[f])
        rab = [int(v) for v in self.fx["reduction_arity_bits"]]
        d.num_reduction_arity_bits = len(rab)
        for i, a in enumerate(rab):
            d.reduction_arity_bits[i] = a
        k["gates"] = (ffi.Gate * len(self.gates))(*[ffi.Gate(*[int(v) for v in g]) for g in self.gates])
        d.num_gates, d.gates = len(self.gates), k["gates"]
        k["lut_lens"] = np.ascontiguousarray(self.fx["lut_lens"] if self.fx["lut_lens"].size else [0], dtype=np.int32)
        k["lut_data"] = np.ascontiguousarray(self.fx["lut_data"] if self.fx["lut_data"].size else [0], dtype=np.uint16)
        k["lookup_rows"] = np.ascontiguousarray(self.fx["lookup_rows"] if self.fx["lookup_rows"].size else [0], dtype=np.int32)
        k["k_is"] = np.ascontiguousarray(self.fx["k_is"], dtype=np.uint64)
        k["cs"] = np.ascontiguousarray(self.fx["constants_sigmas"], dtype=np.uint64)
        d.num_luts = int(self.fx["lut_lens"].size)
        d.lut_lens = k["lut_lens"].ctypes.data_as(C.POINTER(C.c_int32))
        d.lut_data = k["lut_data"].ctypes.data_as(C.POINTER(C.c_uint16))
        d.lookup_rows = k["lookup_rows"].ctypes.data_as(C.POINTER(C.c_int32))
        d.k_is = k["k_is"].ctypes.data_as(C.POINTER(C.c_uint64))
        d.constants_sigmas = k["cs"].ctypes.data_as(C.POINTER(C.c_uint64))
        for i in range(4):
            d.circuit_digest[i] = int(self.circuit_digest[i])
        return d


def from_circuit_data(data, wires, proof_bytes, public_inputs=()):
    """the sections a fixture of this repo's own CircuitData would hold (self-test of the loader only)"""
    d = data.descriptor()
    ncs = d.num_selectors + d.num_lookup_selectors + d.num_constants + d.num_routed_wires
    gates = np.array([[g.kind, g.selector_index, g.group_start, g.group_end, g.num_constraints, g.param0]
                      for g in (d.gates[i] for i in range(d.num_gates))], dtype=np.int32)
    nl = d.num_luts
    lut_lens = np.array([d.lut_lens[i] for i in range(nl)], dtype=np.int32)
    tot = int(lut_lens.sum())
    return {
        "config": np.array([getattr(d, f) for f in CONFIG_FIELDS], dtype=np.int32),
        "reduction_arity_bits": np.array([d.reduction_arity_bits[i] for i in range(d.num_reduction_arity_bits)], dtype=np.int32),
        "gates": gates.ravel(), "gate_ids": np.frombuffer(b"(self-generated)\n" * len(gates), dtype=np.uint8),
        "lut_lens": lut_lens, "lut_data": np.array([d.lut_data[i] for i in range(2 * tot)], dtype=np.uint16),
        "lookup_rows": np.array([d.lookup_rows[i] for i in range(3 * nl)], dtype=np.int32),
        "k_is": np.array([d.k_is[i] for i in range(d.num_routed_wires)], dtype=np.uint64),
        "constants_sigmas": np.ascontiguousarray(data.constants_sigmas[:ncs], dtype=np.uint64).ravel(),
        "circuit_digest": np.array([d.circuit_digest[i] for i in range(4)], dtype=np.uint64),
        "constants_sigmas_cap": np.ascontiguousarray(data.constants_sigmas_cap, dtype=np.uint64).ravel(),
        "wires": np.ascontiguousarray(wires, dtype=np.uint64).ravel(),
        "public_inputs": np.array(list(public_inputs), dtype=np.uint64),
        "proof_bytes": np.frombuffer(proof_bytes, dtype=np.uint8),
    }
