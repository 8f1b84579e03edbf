#!/usr/bin/env python3
"""Writes tests/golden/proof_digests.json: SHA-256 of the proof words, the circuit digest and the
constants/sigmas cap of fixed circuits + witnesses, as produced by the CPU oracle.

These are SELF-GENERATED regression anchors, not reference vectors (no proof bytes of the Rust
plonky2 exist in /root/reference): they catch a simultaneous drift of the oracle and
the CUDA path, which the GPU-vs-oracle parity tests cannot see.  Regenerate only after an intended
protocol change:  python tools/gen_golden_proofs.py"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from tests import circuits, oracle_lib  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


def cases():
    d, w = circuits.tiny_arith()
    yield "tiny_arith", d, w
    d, w, _ = circuits.aes_block()
    yield "aes128_block_fips197_c1", d, w
    d, w, _ = circuits.aes_gcm(13, True)
    yield "aes_gcm_13_bytes_tag", d, w
    d, w, _ = circuits.feistel_poseidon()
    yield "feistel_poseidon_nr32", d, w


def main():
    orc = oracle_lib.load()
    out = {}
    for name, data, wires in cases():
        digest = oracle_lib.set_circuit_digest(orc, data)
        oc = oracle_lib.OracleCircuit(orc, data)
        proof = oc.prove(wires)
        assert oc.verify(proof) == 0
        out[name] = {"n": int(data.n), "proof_words": int(len(proof)), "proof_sha256": sha(proof),
                     "wires_sha256": sha(wires), "constants_sigmas_cap_sha256": sha(oc.cap),
                     "circuit_digest": [int(v) for v in digest]}
        oc.free()
    path = os.path.join(ROOT, "tests", "golden", "proof_digests.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
