"""Shared circuit fixtures for the tests: the reference's own test shapes
(/root/reference/aes-gcm/src/circuit_gcm.rs:737-783 `test_encrypt_op`, circuit_aes.rs:657-726)."""
import functools

import numpy as np

from plonky2_aes_b200.host.circuit_builder import CircuitBuilder, PartialWitness
from plonky2_aes_b200.host.gadgets import native
from plonky2_aes_b200.host.gadgets.aes import (AESStateOps, byte_xor_lut, flatten, from_flat, gf_2_8_mul_lut, sbox_lut,
                                               state_mix_matrix)
from plonky2_aes_b200.host.gadgets.gcm import AesGcmTarget


@functools.lru_cache(maxsize=None)
def tiny_arith():
    """arithmetic-only circuit (no lookups): exercises is_equal/select and n = 8"""
    b = CircuitBuilder()
    x, y = b.add_virtual_target(), b.add_virtual_target()
    z = b.mul(x, y)
    w = b.add(z, x)
    e = b.is_equal(w, y)
    s = b.select(e, x, y)
    b.connect(s, b.add_virtual_target())
    data = b.build()
    pw = PartialWitness()
    pw.set_target(x, 3)
    pw.set_target(y, 5)
    return data, data.generate_witness(pw)


def tiny_arith_padded(degree_bits):
    """the same arithmetic circuit padded with NoopGate rows to n = 2^degree_bits (large-degree prover paths)"""
    b = CircuitBuilder()
    x, y = b.add_virtual_target(), b.add_virtual_target()
    z = b.mul(x, y)
    w = b.add(z, x)
    e = b.is_equal(w, y)
    s = b.select(e, x, y)
    b.connect(s, b.add_virtual_target())
    data = b.build(min_degree_bits=degree_bits)
    pw = PartialWitness()
    pw.set_target(x, 3)
    pw.set_target(y, 5)
    return data, data.generate_witness(pw)


@functools.lru_cache(maxsize=None)
def aes_block(nk=4, nr=10):
    """single AES block encryption circuit, FIPS-197 App. B vector (circuit_aes.rs:619-726)"""
    b = CircuitBuilder()
    ops = AESStateOps(b)
    sb, xo, gf = sbox_lut(b), byte_xor_lut(b), gf_2_8_mul_lut(b)
    key_t = [ops.add_virtual_byte_target(sb) for _ in range(4 * nk)]
    pt_t = [ops.add_virtual_byte_target(sb) for _ in range(16)]
    ct_t = [ops.add_virtual_byte_target(sb) for _ in range(16)]
    mix = state_mix_matrix(b)
    w = ops.key_expansion(nk, nr, xo, sb, key_t)
    out = flatten(ops.encrypt_block(nr, xo, gf, sb, mix, from_flat(pt_t), w))
    for a, c in zip(out, ct_t):
        b.connect(a, c)
    data = b.build()
    # keys of test_encrypt_block_test_vector (circuit_aes.rs:619-655): AES-128 / 192 / 256
    key = bytes.fromhex({4: "2b7e151628aed2a6abf7158809cf4f3c", 6: "8e73b0f7da0e6452c810f32b809079e562f8ead2522c6b7b",
                         8: "603deb1015ca71be2b73aef0857d77811f352c073b6108d72d9810a30914dff4"}[nk])
    pt = bytes.fromhex("3243f6a8885a308d313198a2e0370734")
    ct = bytes(native.flatten_state(native.encrypt_block(list(pt), native.key_expansion(list(key), nk, nr), nr)))
    if nk == 4:
        assert ct == bytes.fromhex("3925841d02dc09fbdc118597196a0b32")      # FIPS-197 App. B
    targets = key_t + pt_t + ct_t
    pw = PartialWitness()
    for t, v in zip(targets, key + pt + ct):
        pw.set_target(t, v)
    return data, data.generate_witness(pw), (targets, key, pt, ct)


@functools.lru_cache(maxsize=None)
def aes_gcm(L=16, tag=True):
    """AesGcmTarget::<4,4,10,L,TAG> with the reference's fixed test pattern
    key=[42;16], nonce=[111;12], pt=[42;L] (circuit_gcm.rs:750-752)"""
    b = CircuitBuilder()
    tg = AesGcmTarget(b, 4, 10, L, tag)
    data = b.build()
    key, nonce, pt = bytes([42] * 16), bytes([111] * 12), bytes([42] * L)
    ct, tagv = native.gcm_encrypt(key, nonce, pt)
    pw = PartialWitness()
    tg.set_targets(pw, key, nonce, pt, ct, tagv)
    return data, data.generate_witness(pw), tg


@functools.lru_cache(maxsize=None)
def public_input_circuit():
    """arithmetic circuit with three registered public inputs (CircuitBuilder::register_public_input):
    exercises PublicInputGate against a non-zero public-input hash computed by in-circuit PoseidonGate rows"""
    b = CircuitBuilder()
    x, y = b.add_virtual_target(), b.add_virtual_target()
    z = b.mul(x, y)
    w = b.add(z, x)
    b.register_public_inputs([x, w, z])
    data = b.build()
    data.test_y_target = y
    pw = PartialWitness()
    pw.set_target(x, 1234567)
    pw.set_target(y, 0xFFFFFFFF00000000)
    slots = data.generate_slots(pw)
    return data, data.generate_witness(pw), data.public_inputs_of(slots)


def gcm_inputs(tg, seed, count):
    """`count` random (key, nonce, pt) with their ct/tag as witness input rows"""
    rows = []
    try:                                   # same ct || tag as gadgets/native.py (tested), much faster
        from cryptography.hazmat.primitives.ciphers.aead import AESGCM
    except ImportError:
        AESGCM = None
    for i in range(count):
        raw = np.random.default_rng(seed + i).bytes(28 + tg.L)
        key, nonce, pt = raw[:16], raw[16:28], raw[28:]
        if AESGCM is not None and tg.nk == 4:
            out = AESGCM(key).encrypt(nonce, pt, None)
            ct, tagv = out[:-16], out[-16:]
        else:
            ct, tagv = native.gcm_encrypt(key, nonce, pt)
        rows.append(tg.input_values(key, nonce, pt, ct, tagv))
    return np.array(rows, dtype=np.uint64)


@functools.lru_cache(maxsize=None)
def feistel_poseidon(nr=32, half=4, key_len=4, seed=7):
    """feistel_poseidon_check (/root/reference/feistel/src/circuit.rs:113-153): NR=32 rounds,
    Poseidon round function, seeded instead of thread_rng inputs"""
    from plonky2_aes_b200.host.gadgets import feistel, poseidon_native
    b = CircuitBuilder()
    st = feistel.add_feistel_state_target(b, 2 * half)
    ks = [b.add_virtual_targets(key_len) for _ in range(nr)]
    out = feistel.feistel_cipher_target(b, st, ks, lambda bb, t: bb.hash_n_to_hash_no_pad(t))
    data = b.build()
    rng = np.random.default_rng(seed)
    P = 0xFFFFFFFF00000001
    state = [int(v) for v in rng.integers(0, P, size=2 * half, dtype=np.uint64)]
    keys = [[int(v) for v in rng.integers(0, P, size=key_len, dtype=np.uint64)] for _ in range(nr)]
    exp = feistel.feistel_cipher(state, keys, poseidon_native.hash_n_to_hash_no_pad)
    assert feistel.feistel_inv_cipher(exp, keys[::-1], poseidon_native.hash_n_to_hash_no_pad) == state
    pw = PartialWitness()
    for t, v in zip(st, state):
        pw.set_target(t, v)
    for kt, kv in zip(ks, keys):
        for t, v in zip(kt, kv):
            pw.set_target(t, v)
    for t, v in zip(out, exp):
        pw.set_target(t, v)
    return data, data.generate_witness(pw), (st, ks, out, state, keys, exp)
