"""CPU: the oracle's restated prover + verifier on the reference's circuit shapes, and the
witness-source vectors the reference's tests pin (FIPS-197, NIST CAVP GCM)."""
import numpy as np
import pytest

from plonky2_aes_b200.host.circuit_builder import PartialWitness
from plonky2_aes_b200.host.gadgets import native
from tests import circuits, oracle_lib

P = 0xFFFFFFFF00000001


def test_native_aes_fips197_vectors():
    # /root/reference/aes-gcm/src/native_aes.rs:165-224 (FIPS-197 App. A first words, App. B block)
    w = native.key_expansion(list(bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")), 4, 10)
    assert bytes(w[4]).hex() == "a0fafe17" and bytes(w[43]).hex() == "b6630ca6"
    w = native.key_expansion(list(bytes.fromhex("8e73b0f7da0e6452c810f32b809079e562f8ead2522c6b7b")), 6, 12)
    assert bytes(w[6]).hex() == "fe0c91f7"
    w = native.key_expansion(list(bytes.fromhex("603deb1015ca71be2b73aef0857d77811f352c073b6108d72d9810a30914dff4")), 8, 14)
    assert bytes(w[8]).hex() == "9ba35411"
    # FIPS-197 §4.2: {57} x {13} = {fe}  (circuit_aes.rs:488-498)
    assert [native.gf_2_8_mul(0x57, v) for v in (1, 2, 4, 8, 0x10, 0x13)] == [0x57, 0xae, 0x47, 0x8e, 0x07, 0xfe]


def test_native_gcm_nist_cavp_vectors():
    # /root/reference/aes-gcm/src/native_gcm.rs:286-330: NIST CAVP AES-128-GCM, 96-bit IV, no AAD
    vecs = [
        # the four vectors the reference holds (CAVP file lines 7, 4417, 8834, 13237)
        ("cf063a34d4a9a76c2c86787d3f96db71", "113b9785971864c83b01c787", "", "", "72ac8493e3a5228b5d130a69d2510e42"),
        ("e98b72a9881a84ca6b76e0f43e68647a", "8b23299fde174053f3d652ba", "28286a321293253c3e0aa2704a278032",
         "5a3c1cf1985dbb8bed818036fdd5ab42", "23c7ab0f952b7091cd324835043b5eb5"),
        ("387218b246c1a8257748b56980e50c94", "dd7e014198672be39f95b69d", "48f5b426baca03064554cc2b30",
         "cdba9e73eaf3d38eceb2b04a8d", "ecf90f4a47c9c626d6fb2c765d201556"),
        ("bfd414a6212958a607a0f5d3ab48471d", "86d8ea0ab8e40dcc481cd0e2",
         "a6b76a066e63392c9443e60272ceaeb9d25c991b0f2e55e2804e168c05ea591a",
         "62171db33193292d930bf6647347652c1ef33316d7feca99d54f1db4fcf513f8", "c28280aa5c6c7a8bd366f28c1cfd1f6e"),
        # two more from the same CAVP set
        ("11754cd72aec309bf52f7687212e8957", "3c819d9a9bed087615030b65", "", "", "250327c674aaf477aef2675748cf6971"),
        ("7fddb57453c241d03efbed3ac44e371c", "ee283a3fc75575e33efd4887", "d5de42b461646c255c87bd2962d3b9a2",
         "2ccda4a5415cb91e135c2a0f78c9b2fd", "b36d1df9b9d5e596f83e8b7f52971cb3"),
    ]
    for key, iv, pt, ct, tag in vecs:
        c, t = native.gcm_encrypt(bytes.fromhex(key), bytes.fromhex(iv), bytes.fromhex(pt))
        assert c.hex() == ct and t.hex() == tag
    from cryptography.hazmat.primitives.ciphers.aead import AESGCM
    rng = np.random.default_rng(7)
    for L in (0, 1, 13, 16, 33, 256):
        k, n, p = rng.bytes(16), rng.bytes(12), rng.bytes(L)
        c, t = native.gcm_encrypt(k, n, p)
        assert c + t == AESGCM(k).encrypt(n, p, None)


def test_tiny_circuit_prove_verify(oracle):
    data, wires = circuits.tiny_arith()
    oc = oracle_lib.OracleCircuit(oracle, data)
    proof = oc.prove(wires)
    assert oc.verify(proof) == 0
    bad = proof.copy(); bad[len(bad) // 2] ^= 1
    assert oc.verify(bad) != 0
    w2 = wires.copy(); w2[3, 0] = (int(w2[3, 0]) + 1) % P          # break an arithmetic gate
    assert oc.verify(oc.prove(w2)) == -20
    oc.free()


def test_aes_block_circuit_prove_verify(oracle):
    """C1: the reference's single-block test circuit with the FIPS-197 App. B witness."""
    data, wires, (targets, key, pt, ct) = circuits.aes_block()
    assert data.n == 8192
    oc = oracle_lib.OracleCircuit(oracle, data)
    proof, tr, zs, qc = oc.prove(wires, debug=True)
    assert oc.verify(proof) == 0
    assert tr.pow_witness < (1 << 24)
    assert int(zs[0, 0]) == 1 and int(zs[1, 0]) == 1          # Z(1) = 1
    # wrong ciphertext: the witness generator refuses (prove() -> Err, circuit_aes.rs:403-405)
    pw = PartialWitness()
    badct = bytes([ct[0] ^ 1]) + ct[1:]
    for t, v in zip(targets, key + pt + badct):
        pw.set_target(t, v)
    with pytest.raises(ValueError):
        data.generate_witness(pw)
    # non-byte input (test_assert_byte, circuit_aes.rs:385-411)
    pw = PartialWitness()
    for t, v in zip(targets, [256] + list(key[1:]) + list(pt) + list(ct)):
        pw.set_target(t, v)
    with pytest.raises(ValueError):
        data.generate_witness(pw)
    # a forged witness cell is caught by the verifier
    w2 = wires.copy(); w2[1, 5] ^= 1
    assert oc.verify(oc.prove(w2)) != 0
    oc.free()


def test_aes_gcm_tag_circuit(oracle):
    """AesGcmTarget<4,4,10,13,true> — the reference's TAG=true test shape (circuit_gcm.rs:698)."""
    data, wires, tg = circuits.aes_gcm(13, True)
    oc = oracle_lib.OracleCircuit(oracle, data)
    assert oc.verify(oc.prove(wires)) == 0
    # batch witness generation agrees with the single path
    vals = circuits.gcm_inputs(tg, 99, 3)
    many = data.generate_witnesses(tg.input_targets(), vals)
    pw = PartialWitness()
    for t, v in zip(tg.input_targets(), vals[1]):
        pw.set_target(t, int(v))
    assert np.array_equal(many[1], data.generate_witness(pw))
    assert oc.verify(oc.prove(many[2])) == 0
    oc.free()


def test_feistel_poseidon_circuit(oracle):
    """config 3 (Feistel half): 32 PoseidonGate rows — feistel_poseidon_check,
    /root/reference/feistel/src/circuit.rs:113-153, and the native round trip lib.rs:98-120."""
    data, wires, (st, ks, out, state, keys, exp) = circuits.feistel_poseidon()
    assert data.num_selectors == 2 and data.num_gate_constraints == 123
    oc = oracle_lib.OracleCircuit(oracle, data)
    assert oc.verify(oc.prove(wires)) == 0
    # wrong expected output is refused by the witness generator (prove() -> Err)
    pw = PartialWitness()
    for t, v in zip(st, state):
        pw.set_target(t, v)
    for kt, kv in zip(ks, keys):
        for t, v in zip(kt, kv):
            pw.set_target(t, v)
    pw.set_target(out[0], (exp[0] + 1) % P)
    with pytest.raises(ValueError):
        data.generate_witness(pw)
    # a forged S-box wire inside a PoseidonGate row is caught by the verifier
    w2 = wires.copy(); w2[70, 3] ^= 1
    assert oc.verify(oc.prove(w2)) == -20
    oc.free()


def test_proof_views_and_byte_round_trip(oracle):
    from plonky2_aes_b200.host.proof import Proof
    data, wires, _ = circuits.aes_gcm(13, True)
    oc = oracle_lib.OracleCircuit(oracle, data)
    words, tr, _, _ = oc.prove(wires, debug=True)
    pr = Proof(words, data.descriptor())
    assert pr.pow_witness == tr.pow_witness
    assert pr["wires_cap"].size == 64 and pr["openings.wires"].size == 270
    assert int(pr["fri.query[0].initial[1].path_len"][0]) == data.degree_bits + 3 - 4
    raw = pr.to_bytes()                                  # p2g_proof_to_bytes (C ABI, host-only)
    n_paths = data.config.fri_config.num_query_rounds * (4 + len(data.reduction_arity_bits))
    assert len(raw) == 8 * (len(words) - n_paths) + n_paths + 8        # + the public-input count (u64)
    # independent restatement of upstream's write order (util/serialization/mod.rs): caps; openings with
    # lookup_zs / lookup_zs_next right after plonk_zs_next; FRI caps; queries (u8 path lengths); final poly;
    # pow witness; number of public inputs; public inputs
    names = list(pr.segments)
    op = [s for s in names if s.startswith("openings.")]
    assert [s.split(".")[1] for s in op] == ["constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next",
                                              "partial_products", "quotient_polys", "lookup_zs", "lookup_zs_next"]
    order = names[:3] + op[:5] + op[7:9] + op[5:7] + [s for s in names[3:] if not s.startswith("openings.")]
    exp = bytearray()
    for s in order:
        pos, cnt, kind = pr.segments[s]
        if s == "public_inputs":
            exp += (cnt).to_bytes(8, "little")
        exp += bytes([int(words[pos])]) if kind == "len" else words[pos:pos + cnt].astype("<u8").tobytes()
    assert raw == bytes(exp)
    back = Proof.from_bytes(raw, data.descriptor())
    assert np.array_equal(back.words, words) and oc.verify(back.words) == 0
    for bad in (raw[:-1], raw + b"\0", raw[:200] + b"\xff" * 8 + raw[208:]):       # truncated, trailing, non-canonical
        with pytest.raises(ValueError):
            Proof.from_bytes(bad, data.descriptor())
    oc.free()


def test_golden_proof_digests(oracle):
    """Regression anchors (tests/golden/proof_digests.json, written by tools/gen_golden_proofs.py): the
    oracle reproduces the committed circuit digests and proof hashes bit for bit.  Self-generated, not
    reference vectors -- they pin today's behaviour of the restated protocol against silent drift."""
    import hashlib
    import json
    import os
    from tools import gen_golden_proofs as g
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "proof_digests.json")))
    seen = set()
    for name, data, wires in g.cases():
        ref = gold[name]
        digest = oracle_lib.set_circuit_digest(oracle, data)
        assert [int(v) for v in digest] == ref["circuit_digest"], name
        assert g.sha(wires) == ref["wires_sha256"], name
        oc = oracle_lib.OracleCircuit(oracle, data)
        assert g.sha(oc.cap) == ref["constants_sigmas_cap_sha256"], name
        proof = oc.prove(wires)
        assert len(proof) == ref["proof_words"] and g.sha(proof) == ref["proof_sha256"], name
        oc.free()
        seen.add(name)
    assert seen == set(gold)


def test_public_inputs_prove_verify(oracle):
    """register_public_input: the PublicInputGate ties wires 0..4 to the in-circuit hash of the public
    inputs; the verifier recomputes it from the proof tail."""
    data, wires, pi = circuits.public_input_circuit()
    oracle_lib.set_circuit_digest(oracle, data)
    oc = oracle_lib.OracleCircuit(oracle, data)
    proof = oc.prove(wires, public_inputs=pi)
    assert list(proof[-3:]) == list(pi)
    assert oc.verify(proof) == 0
    bad = proof.copy(); bad[-1] ^= 1
    assert oc.verify(bad) != 0
    oc.free()


def test_aes192_block_circuit(oracle):
    """NK=6, NR=12 of test_encrypt_block_test_vector (/root/reference/aes-gcm/src/circuit_aes.rs:631-640)"""
    data, wires, _ = circuits.aes_block(6, 12)
    oracle_lib.set_circuit_digest(oracle, data)
    oc = oracle_lib.OracleCircuit(oracle, data)
    assert oc.verify(oc.prove(wires)) == 0
    oc.free()
