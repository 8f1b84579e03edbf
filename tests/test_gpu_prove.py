"""GPU parity of the whole prove() path through the C ABI (p2g_prove) against the CPU oracle:
identical transcript, identical Z / partial-product / lookup polynomials, identical quotient
chunks, bit-identical proof words, and every GPU proof accepted by the restated verifier."""
import ctypes as C

import numpy as np
import pytest

from plonky2_aes_b200.host import ffi
from tests import circuits, oracle_lib

pytestmark = pytest.mark.gpu
P = 0xFFFFFFFF00000001


def _gpu_debug(ctx, data, wires):
    lib = ctx.lib
    ctx.check(lib.p2g_set_timing(ctx.handle, 3))
    proof = data.prove_wires(wires)
    tr = ffi.Transcript()
    ctx.check(lib.p2g_last_transcript(ctx.handle, C.byref(tr)))
    d = data.descriptor()
    nlp = 0 if d.num_luts == 0 else -(-(d.num_routed_wires // 2) // (d.quotient_degree_factor - 1)) + 1
    zs = np.empty((d.num_challenges * (1 + d.num_partial_products + nlp), data.n), dtype=np.uint64)
    qc = np.empty((d.num_challenges * d.quotient_degree_factor, data.n), dtype=np.uint64)
    ctx.check(lib.p2g_last_zs_values(ctx.handle, zs.ctypes.data))
    ctx.check(lib.p2g_last_quotient_chunks(ctx.handle, qc.ctypes.data))
    ctx.check(lib.p2g_set_timing(ctx.handle, 0))
    return proof, tr, zs, qc


def _check(ctx, oracle, data, wires):
    data.load(ctx)
    oc = oracle_lib.OracleCircuit(oracle, data)
    assert np.array_equal(oc.cap, data.constants_sigmas_cap)
    o_proof, o_tr, o_zs, o_qc = oc.prove(wires, debug=True)
    g_proof, g_tr, g_zs, g_qc = _gpu_debug(ctx, data, wires)
    for f in ("betas", "gammas", "deltas"):
        assert list(getattr(g_tr, f)) == list(getattr(o_tr, f)), f
    assert np.array_equal(g_zs, o_zs), "Z / partial products / lookup polynomials"
    assert list(g_tr.alphas) == list(o_tr.alphas)
    assert np.array_equal(g_qc, o_qc), "quotient chunks"
    for f in ("zeta", "fri_alpha", "fri_betas", "query_indices"):
        assert list(getattr(g_tr, f)) == list(getattr(o_tr, f)), f
    assert g_tr.pow_witness == o_tr.pow_witness
    assert len(g_proof) == len(o_proof)
    assert np.array_equal(g_proof, o_proof), "proof words"
    assert oc.verify(g_proof) == 0
    return oc


def test_tiny_circuit(gpu_ctx, oracle):
    data, wires = circuits.tiny_arith()
    oc = _check(gpu_ctx, oracle, data, wires)
    w2 = wires.copy(); w2[3, 0] = (int(w2[3, 0]) + 1) % P
    assert oc.verify(data.prove_wires(w2)) == -20        # unsatisfied witness -> rejected proof
    oc.free()


def test_degree_2_18(gpu_ctx, oracle):
    """n = 2^18 (LDE of 2^21 leaves): the pre-folded NTT inside the whole prover, every stage against the oracle"""
    data, wires = circuits.tiny_arith_padded(18)
    assert data.degree_bits == 18
    oc = _check(gpu_ctx, oracle, data, wires)
    oc.free()
    gpu_ctx.check(gpu_ctx.lib.p2g_circuit_free(gpu_ctx.handle, data._gpu_circuit))


def test_aes_block_c1(gpu_ctx, oracle):
    data, wires, _ = circuits.aes_block()
    _check(gpu_ctx, oracle, data, wires).free()


def test_aes_gcm_tag_small(gpu_ctx, oracle):
    data, wires, tg = circuits.aes_gcm(13, True)
    oc = _check(gpu_ctx, oracle, data, wires)
    many = data.generate_witnesses(tg.input_targets(), circuits.gcm_inputs(tg, 5, 2))
    for w in many:
        assert oc.verify(data.prove_wires(w)) == 0
    oc.free()


def test_aes_gcm_256_tag_c2(gpu_ctx, oracle):
    """BASELINE config 2: AES-GCM-128, 256-byte plaintext, with tag; n = 2^15."""
    data, wires, tg = circuits.aes_gcm(256, True)
    assert data.n == 1 << 15
    _check(gpu_ctx, oracle, data, wires).free()


def test_aes256_gcm_and_no_tag_variants(gpu_ctx, oracle):
    """AES-256 key schedule (NK=8, NR=14) and the TAG=false alias (`AesGcm128Target`, lib.rs:19)."""
    from plonky2_aes_b200.host.circuit_builder import CircuitBuilder, PartialWitness
    from plonky2_aes_b200.host.gadgets import native
    from plonky2_aes_b200.host.gadgets.gcm import AesGcmTarget
    for nk, nr, L, tag in ((8, 14, 17, True), (4, 10, 42, False)):
        b = CircuitBuilder()
        tg = AesGcmTarget(b, nk, nr, L, tag)
        data = b.build()
        key, nonce, pt = bytes(range(4 * nk)), bytes([111] * 12), bytes([231] * L)
        ct, tagv = native.gcm_encrypt(key, nonce, pt, nk=nk, nr=nr)
        pw = PartialWitness()
        tg.set_targets(pw, key, nonce, pt, ct, tagv)
        _check(gpu_ctx, oracle, data, data.generate_witness(pw)).free()


def test_feistel_poseidon_gate(gpu_ctx, oracle):
    """config 3 (Feistel half): PoseidonGate constraints in the quotient kernel, two selector groups."""
    data, wires, _ = circuits.feistel_poseidon()
    _check(gpu_ctx, oracle, data, wires).free()


def test_stage_helpers(gpu_ctx, oracle):
    rng = np.random.default_rng(3)
    # proof-of-work grind: lowest nonce
    st = rng.integers(0, P, size=12, dtype=np.uint64)
    nonce = C.c_uint64()
    gpu_ctx.check(gpu_ctx.lib.p2g_pow_grind(gpu_ctx.handle, st.ctypes.data, 5, 12, C.byref(nonce)))
    best = None
    for cand in range(nonce.value + 1):
        s = st.copy(); s[5] = cand
        if int(oracle.poseidon(s)[7]) >> 52 == 0:
            best = cand
            break
    assert best == nonce.value
    # FRI fold against the coefficient-domain definition
    log_len, ab = 8, 4
    ln = 1 << log_len
    coeffs = rng.integers(0, P, size=(ln, 2), dtype=np.uint64)
    shift = 7
    vals = np.stack([oracle.coset_fft(coeffs[:, 0], shift), oracle.coset_fft(coeffs[:, 1], shift)], axis=1)
    rev = np.array([int(format(i, f"0{log_len}b")[::-1], 2) for i in range(ln)])
    vals_br = np.ascontiguousarray(vals[rev])
    beta = rng.integers(0, P, size=2, dtype=np.uint64)
    out = np.empty((ln >> ab, 2), dtype=np.uint64)
    gpu_ctx.check(gpu_ctx.lib.p2g_fri_fold(gpu_ctx.handle, vals_br.ctypes.data, log_len, ab, shift, beta.ctypes.data, out.ctypes.data))

    def emul(a, b):
        return ((a[0] * b[0] + 7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
    folded = []
    b = (int(beta[0]), int(beta[1]))
    for k in range(ln >> ab):
        acc = (0, 0)
        for i in reversed(range(1 << ab)):
            acc = emul(acc, b)
            acc = ((acc[0] + int(coeffs[16 * k + i, 0])) % P, (acc[1] + int(coeffs[16 * k + i, 1])) % P)
        folded.append(acc)
    folded = np.array(folded, dtype=np.uint64)
    nshift = pow(shift, 1 << ab, P)
    nv = np.stack([oracle.coset_fft(folded[:, 0], nshift), oracle.coset_fft(folded[:, 1], nshift)], axis=1)
    rev2 = np.array([int(format(i, f"0{log_len - ab}b")[::-1], 2) for i in range(ln >> ab)])
    assert np.array_equal(out, nv[rev2])


def test_error_behaviour(gpu_ctx, oracle):
    """argument errors come back as codes (no crash, no fallback): the C ABI's contract."""
    lib = gpu_ctx.lib
    h = C.c_void_p()
    cols = np.zeros((1, 8), dtype=np.uint64)
    assert lib.p2g_commit_from_values(gpu_ctx.handle, cols.ctypes.data, 0, 3, 3, 4, C.byref(h), None) == -2   # no columns
    assert lib.p2g_commit_from_values(gpu_ctx.handle, cols.ctypes.data, 1, 3, 3, 9, C.byref(h), None) == -2   # cap above the tree
    assert lib.p2g_commit_from_values(gpu_ctx.handle, None, 1, 3, 3, 4, C.byref(h), None) == -2
    data, wires = circuits.tiny_arith()
    data.load(gpu_ctx)
    # a configuration the prover does not implement is refused at load time
    desc = data.descriptor()
    desc.rate_bits = 2
    assert lib.p2g_circuit_load(gpu_ctx.handle, C.byref(desc), C.byref(h), None) == -2
    # proof buffer too small
    out = np.zeros(16, dtype=np.uint64)
    got = C.c_size_t()
    assert lib.p2g_prove(gpu_ctx.handle, data._gpu_circuit, wires.ctypes.data, None, out.ctypes.data, out.size, C.byref(got)) == -2
    # the context is still usable afterwards
    assert oracle_lib.OracleCircuit(oracle, data).verify(data.prove_wires(wires)) == 0


def test_prove_slots_device_wire_fill(gpu_ctx, oracle):
    """p2g_prove_slots: the wire matrix gathered on the device from one value per partition
    (PartitionWitness::full_witness, iop/witness.rs) equals the host fill, and the proof is
    bit-identical to p2g_prove on the host matrix.  AES-GCM (LUT multiplicities, constant cells)
    and Feistel (PoseidonGate internal wires)."""
    lib = gpu_ctx.lib
    data, wires, tg = circuits.aes_gcm(13, True)
    data.load(gpu_ctx)
    vals = circuits.gcm_inputs(tg, 11, 3)
    host = data.generate_witnesses(tg.input_targets(), vals)
    slots = data.generate_slots_many(tg.input_targets(), vals)
    assert slots.shape == (3, data.ext_slots)
    wmap = data.load_wire_map()
    oc = oracle_lib.OracleCircuit(oracle, data)
    for i in range(3):
        filled = np.empty_like(host[i])
        gpu_ctx.check(lib.p2g_wmap_fill(gpu_ctx.handle, data._gpu_circuit, wmap, slots[i].ctypes.data, filled.ctypes.data))
        assert np.array_equal(filled, host[i])
        p_slots = data.prove_slots(slots[i], wmap=wmap)
        assert np.array_equal(p_slots, data.prove_wires(host[i]))
        assert oc.verify(p_slots) == 0
    oc.free()
    # bad arguments: a map entry past the slot vector
    bad = np.full((135, data.n), data.ext_slots, dtype=np.int32)
    h = C.c_void_p()
    assert lib.p2g_wmap_load(gpu_ctx.handle, data._gpu_circuit, bad.ctypes.data, data.ext_slots, None, None, 0, C.byref(h)) == -2
    gpu_ctx.check(lib.p2g_wmap_free(gpu_ctx.handle, wmap))

    fdata, fwires, (st, ks, out, state, keys, exp) = circuits.feistel_poseidon()
    from plonky2_aes_b200.host.circuit_builder import PartialWitness
    fdata.load(gpu_ctx)
    pw = PartialWitness()
    for t, v in zip(st, state):
        pw.set_target(t, v)
    for kt, kv in zip(ks, keys):
        for t, v in zip(kt, kv):
            pw.set_target(t, v)
    for t, v in zip(out, exp):
        pw.set_target(t, v)
    assert np.array_equal(fdata.prove(pw), fdata.prove_wires(fwires))      # prove(pw) goes through the slot path


def test_batch_prover_proofs_in_flight(gpu_ctx, oracle):
    """Several proofs in flight on one GPU (one context + host thread each, the bench.py / config-5
    scheme): proofs are identical to the one-at-a-time proofs, for host-filled wires and for the
    device wire fill, also when the same batch is proved twice (pool reuse)."""
    from plonky2_aes_b200.host.polynomial_batch import Context
    from plonky2_aes_b200.host.sharding import BatchProver
    data, _, tg = circuits.aes_gcm(13, True)
    data.load(gpu_ctx)
    vals = circuits.gcm_inputs(tg, 31, 7)
    host = data.generate_witnesses(tg.input_targets(), vals)
    slots = data.generate_slots_many(tg.input_targets(), vals)
    seq = [data.prove_wires(w) for w in host]
    extra = [Context(0) for _ in range(3)]
    bp = BatchProver(data, [gpu_ctx] + extra)
    for _ in range(2):
        par = bp.prove_many([w for w in host])
        par_slots = bp.prove_many([s for s in slots], slots=True)
        for a, b, c in zip(seq, par, par_slots):
            assert np.array_equal(a, b) and np.array_equal(a, c)
    oc = oracle_lib.OracleCircuit(oracle, data)
    assert all(oc.verify(p) == 0 for p in par_slots)
    oc.free()
    for c in extra:
        c.close()


def test_aes_gcm_1024_bytes_n65536(gpu_ctx, oracle):
    """A larger instance of the same circuit family: 64 AES blocks, n = 2^16 (LDE domain 2^19, 2^14-point
    NTT chunks with a 4-way fold, final FRI polynomial of 16 coefficients)."""
    data, wires, _ = circuits.aes_gcm(1024, True)
    assert data.n == 1 << 16
    _check(gpu_ctx, oracle, data, wires).free()


def test_repeated_proofs_do_not_leak(gpu_ctx, oracle):
    """Temporaries of a proof are owned by a scope guard (csrc/prover.cu, Scratch) and go back to the
    context's pool on every return path: device memory outside the pool stays flat over repeated
    proofs, including proofs of a witness that violates the circuit (the prover still answers, and
    the restated verifier rejects that proof -- upstream's `data.verify` would too)."""
    import torch
    data, wires, _ = circuits.aes_gcm(13, True)
    data.load(gpu_ctx)
    oc = oracle_lib.OracleCircuit(oracle, data)
    good = data.prove_wires(wires)
    bad = wires.copy()
    bad[0, 5] = (int(bad[0, 5]) + 1) % P
    free = []
    for it in range(10):
        proof = data.prove_wires(bad if it % 2 else wires)
        assert oc.verify(proof) == (-20 if it % 2 else 0)
        gpu_ctx.sync()
        torch.cuda.synchronize()
        free.append(torch.cuda.mem_get_info()[0])
    assert free[-1] >= free[2], f"device memory shrinks by {(free[2] - free[-1]) / 7 / 1e6:.1f} MB per proof"
    assert np.array_equal(data.prove_wires(wires), good)
    oc.free()


def test_golden_proof_digests_gpu(gpu_ctx):
    """The CUDA path reproduces the committed golden circuit digests and proof hashes
    (tests/golden/proof_digests.json) without consulting the oracle."""
    import json
    import os
    from tools import gen_golden_proofs as g
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "proof_digests.json")))
    for name, data, wires in g.cases():
        ref = gold[name]
        data.load(gpu_ctx)
        assert [int(v) for v in data.circuit_digest] == ref["circuit_digest"], name
        assert g.sha(data.constants_sigmas_cap) == ref["constants_sigmas_cap_sha256"], name
        proof = data.prove_wires(wires)
        assert len(proof) == ref["proof_words"] and g.sha(proof) == ref["proof_sha256"], name


def test_aes192_block(gpu_ctx, oracle):
    """AES-192 (NK=6, NR=12) single-block circuit, the middle case of test_encrypt_block_test_vector
    (/root/reference/aes-gcm/src/circuit_aes.rs:619-655)."""
    data, wires, _ = circuits.aes_block(6, 12)
    _check(gpu_ctx, oracle, data, wires).free()


def test_public_inputs(gpu_ctx, oracle):
    """register_public_input: PublicInputGate against a non-zero hash, public inputs at the proof tail."""
    data, wires, pi = circuits.public_input_circuit()
    data.load(gpu_ctx)
    oc = oracle_lib.OracleCircuit(oracle, data)
    o_proof = oc.prove(wires, public_inputs=pi)
    g_proof = data.prove_wires(wires, pi)
    assert np.array_equal(g_proof, o_proof)
    assert list(g_proof[-len(pi):]) == list(pi)
    assert oc.verify(g_proof) == 0
    bad = pi.copy(); bad[0] += 1                          # proof for other public inputs: rejected
    assert oc.verify(data.prove_wires(wires, bad)) != 0
    assert np.array_equal(data.prove_slots(data.generate_slots(_pw_public(data))), o_proof)
    oc.free()


def _pw_public(data):
    from plonky2_aes_b200.host.circuit_builder import PartialWitness
    pw = PartialWitness()
    x, w, z = data.public_input_targets
    pw.set_target(x, 1234567)
    pw.set_target(z, (1234567 * 0xFFFFFFFF00000000) % P)     # consistent with y below: no conflict
    pw.set_target(data.test_y_target, 0xFFFFFFFF00000000)
    return pw


def test_circuit_load_error_paths_do_not_leak(gpu_ctx):
    """p2g_circuit_load refusing a descriptor (bad lookup rows: detected after the preprocessed commitment
    was built) must give every device buffer back."""
    import torch
    data, _, _ = circuits.aes_gcm(13, True)
    data.load(gpu_ctx)
    lib = gpu_ctx.lib
    d = data.descriptor()
    rows = np.array(data._lookup_rows_arr, dtype=np.int32).copy()
    rows[2] = data.n                                       # first_lut_gate + 1 >= n
    d.lookup_rows = rows.ctypes.data_as(C.POINTER(C.c_int32))
    h = C.c_void_p()
    assert lib.p2g_circuit_load(gpu_ctx.handle, C.byref(d), C.byref(h), None) == -2     # warm the pool
    gpu_ctx.sync(); torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(5):
        assert lib.p2g_circuit_load(gpu_ctx.handle, C.byref(d), C.byref(h), None) == -2
    gpu_ctx.sync(); torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free1 >= free0 - (1 << 20), (free0, free1)
    d2 = data.descriptor(); d2.pow_bits = 0
    assert lib.p2g_circuit_load(gpu_ctx.handle, C.byref(d2), C.byref(h), None) == -2
    d3 = data.descriptor(); d3.degree_bits = 21
    assert lib.p2g_circuit_load(gpu_ctx.handle, C.byref(d3), C.byref(h), None) == -2


def test_device_witness_generation(gpu_ctx, oracle):
    """p2g_wprog_load / p2g_wprog_generate / p2g_prove_inputs: the level-scheduled generator program on the
    device gives the host generators' extended slot vector bit for bit (AES-GCM: arithmetic, lookup and
    equality generators + multiplicities; Feistel: PoseidonGate generator), the proof from input values
    equals the proof from the host-filled wire matrix, and generator failures come back as P2W_E_* codes."""
    lib = gpu_ctx.lib
    data, wires, tg = circuits.aes_gcm(13, True)
    data.load(gpu_ctx)
    targets = tg.input_targets()
    wp = data.load_witness_program(gpu_ctx, targets)
    assert lib.p2g_wprog_levels(wp) > 100
    vals = circuits.gcm_inputs(tg, 11, 5)
    host = data.generate_slots_many(targets, vals)
    dev = data.generate_slots_device(gpu_ctx, wp, vals)
    assert np.array_equal(dev, host)
    oc = oracle_lib.OracleCircuit(oracle, data)
    w = data.generate_witnesses(targets, vals[:1])[0]
    p_host = data.prove_wires(w)
    p_dev = data.prove_inputs(vals[0], wp)
    assert np.array_equal(p_dev, p_host) and oc.verify(p_dev) == 0
    # batch form with the witnesses kept in HBM (p2g_wprog_generate_dev -> p2g_prove_slots_dev)
    import torch
    ext_dev = torch.empty((5, data.ext_slots), dtype=torch.int64, device="cuda")
    flags = torch.ones(5, dtype=torch.int32, device="cuda")
    gpu_ctx.check(lib.p2g_wprog_generate_dev(gpu_ctx.handle, wp, np.ascontiguousarray(vals).ctypes.data, 5, ext_dev.data_ptr(), flags.data_ptr()))
    words, got = np.empty(data.proof_words, dtype=np.uint64), C.c_size_t()
    gpu_ctx.check(lib.p2g_prove_slots_dev(gpu_ctx.handle, data._gpu_circuit, data._wmap, ext_dev[0].data_ptr(), None, words.ctypes.data,
                                          words.size, C.byref(got)))
    assert np.array_equal(words, p_host) and int(flags.cpu().abs().sum()) == 0
    assert np.array_equal(ext_dev.cpu().numpy().view(np.uint64), host)
    bad = vals[:1].copy(); bad[0, -1] ^= 1                     # wrong tag byte: the computed tag disagrees
    with pytest.raises(ValueError, match="set twice"):
        data.generate_slots_device(gpu_ctx, wp, bad)
    with pytest.raises(ValueError, match="set twice"):
        data.prove_inputs(bad[0], wp)
    bad = vals[:1].copy(); bad[0, 0] = 300                     # a "byte" outside the byte table
    with pytest.raises(ValueError, match="lookup"):
        data.generate_slots_device(gpu_ctx, wp, bad)
    gpu_ctx.check(lib.p2g_wprog_free(gpu_ctx.handle, wp))
    oc.free()
    # PoseidonGate generator + preset outputs that are checked, not written
    data, wires, (st, ks, out, state, keys, exp) = circuits.feistel_poseidon()
    data.load(gpu_ctx)
    targets = list(st) + [t for k in ks for t in k] + list(out)
    v = np.array([list(state) + [x for k in keys for x in k] + list(exp)], dtype=np.uint64)
    wp = data.load_witness_program(gpu_ctx, targets)
    assert np.array_equal(data.generate_slots_device(gpu_ctx, wp, v), data.generate_slots_many(targets, v))
    v[0, -1] = (int(v[0, -1]) + 1) % P
    with pytest.raises(ValueError, match="set twice"):
        data.generate_slots_device(gpu_ctx, wp, v)
    gpu_ctx.check(lib.p2g_wprog_free(gpu_ctx.handle, wp))
    # a program whose inputs are incomplete is refused at load (the host generators would stop with P2W_E_UNSET)
    data, _, tg = circuits.aes_gcm(13, True)
    h = C.c_void_p()
    data._program()
    slots = np.array([data._slot(t) for t in tg.input_targets()[:-20]], dtype=np.int32)
    rc = lib.p2g_wprog_load(gpu_ctx.handle, C.byref(data._wdesc), slots.ctypes.data, len(slots), C.byref(h))
    assert rc in (0, -2)          # without the ciphertext / tag inputs the program computes them instead of checking
    if rc == 0:
        lib.p2g_wprog_free(gpu_ctx.handle, h)
    slots = slots[:10]            # without the plaintext the AES rounds read partitions nobody sets
    assert lib.p2g_wprog_load(gpu_ctx.handle, C.byref(data._wdesc), slots.ctypes.data, len(slots), C.byref(h)) == -2


@pytest.mark.parametrize("world", [2, 4, 8])
def test_coset_sharded_proof_equals_single_gpu_proof(gpu_ctx, oracle, world):
    """p2g_prove_sharded: one proof split by coset over `world` ranks (here host threads on one GPU, each with its
    own context and exchange buffers) is bit-identical to the unsharded proof on every rank -- commitments of the
    owned leaf blocks only, quotient evaluated per coset, all-gathers of cap entries / quotient interpolants /
    last FRI layer / query records.  AES-GCM (lookups), Feistel (PoseidonGate) and the public-input circuit."""
    from plonky2_aes_b200.host.polynomial_batch import Context
    from plonky2_aes_b200.host.sharding import ThreadedShards
    cases = [circuits.aes_gcm(13, True)[:2] + (None,), circuits.feistel_poseidon()[:2] + (None,)]
    if world <= 4:
        cases.append(circuits.aes_block()[:2] + (None,))
    ctxs = [Context(0) for _ in range(world)]
    for data, wires, pi in cases:
        data.load(gpu_ctx)
        ref = data.prove_wires(wires, pi)
        handles = [data.load_handle(c) for c in ctxs]
        proofs = ThreadedShards(world).prove(ctxs, handles, data.proof_words, wires, pi)
        for r, p in enumerate(proofs):
            assert np.array_equal(p, ref), (world, r)
        for c, h in zip(ctxs, handles):
            c.check(c.lib.p2g_circuit_free(c.handle, h))
    for c in ctxs:
        c.close()


def test_prove_batch_entry_point(gpu_ctx, oracle):
    """p2g_prove_batch: the C-level batch (host threads inside the library, proof i on context i mod n_ctx) returns the
    proofs p2g_prove returns one at a time"""
    from plonky2_aes_b200.host.polynomial_batch import Context
    data, _, tg = circuits.aes_gcm(13, True)
    data.load(gpu_ctx)
    wires = data.generate_witnesses(tg.input_targets(), circuits.gcm_inputs(tg, 21, 5))
    ctxs = [gpu_ctx, Context(0), Context(0)]
    handles = [data._gpu_circuit] + [data.load_handle(c) for c in ctxs[1:]]
    words = data.proof_words
    out = np.zeros((5, words), dtype=np.uint64)
    status = np.full(5, 99, dtype=np.int32)
    vp = C.c_void_p
    rc = gpu_ctx.lib.p2g_prove_batch((vp * 3)(*[c.handle for c in ctxs]), (vp * 3)(*handles), 3,
                                     (vp * 5)(*[w.ctypes.data for w in wires]), None, 5,
                                     (vp * 5)(*[o.ctypes.data for o in out]), words, status.ctypes.data)
    assert rc == 0 and not status.any()
    for i in range(5):
        assert np.array_equal(out[i], data.prove_wires(wires[i]))
    for c, h in zip(ctxs[1:], handles[1:]):
        c.check(c.lib.p2g_circuit_free(c.handle, h)); c.close()


def test_guard_bands_see_no_out_of_bounds_write(oracle):
    """compute-sanitizer is closed on the GPU pool, so the library carries its own out-of-bounds-WRITE detector:
    with P2G_CANARY=1 every device block gets a guard band that is checked on release.  Run over the shapes with
    the most irregular indexing: tiny / odd commitments, a lookup circuit, the PoseidonGate circuit, public inputs,
    the device witness generator and a coset-sharded proof."""
    import os
    from plonky2_aes_b200.host.polynomial_batch import Context, PolynomialBatch
    from plonky2_aes_b200.host.sharding import ThreadedShards
    os.environ["P2G_CANARY"] = "1"
    try:
        ctxs = [Context(0) for _ in range(3)]
    finally:
        del os.environ["P2G_CANARY"]
    ctx = ctxs[0]
    rng = np.random.default_rng(1)
    # (2, 17): the pre-folded NTT (outer kernel, scratch for the natural-order inverse); (1, 8, 0): a tree folded to one root
    for ncols, log_n, cap in ((1, 1, 0), (5, 1, 3), (3, 2, 1), (9, 6, 4), (135, 10, 4), (3, 14, 4), (2, 17, 4), (1, 8, 0)):
        PolynomialBatch.from_values(ctx, rng.integers(0, P, size=(ncols, 1 << log_n), dtype=np.uint64), 3, cap).free()
    for data, wires, pi in (circuits.tiny_arith() + (None,), circuits.aes_gcm(13, True)[:2] + (None,),
                            circuits.feistel_poseidon()[:2] + (None,), circuits.public_input_circuit()):
        data.load(ctx)
        data.prove_wires(wires, pi)
    data, wires, tg = circuits.aes_gcm(13, True)
    data.load(ctx)
    wp = data.load_witness_program(ctx, tg.input_targets())
    data._wmap = None
    data.prove_inputs(circuits.gcm_inputs(tg, 3, 1)[0], wp)
    handles = [data._gpu_circuit] + [data.load_handle(c) for c in ctxs[1:]]
    ThreadedShards(2).prove(ctxs[:2], handles[:2], data.proof_words, wires)
    checked = 0
    for c in ctxs:
        n, bad = C.c_uint64(), C.c_uint64()
        c.check(c.lib.p2g_debug_canary(c.handle, C.byref(n), C.byref(bad)))
        assert bad.value == 0
        checked += n.value
    assert checked > 100
    for d in (circuits.tiny_arith()[0], circuits.aes_gcm(13, True)[0], circuits.feistel_poseidon()[0], circuits.public_input_circuit()[0]):
        d._gpu_circuit = None; d.ctx = None; d._wmap = None      # handles of the contexts closed below
    for c in ctxs:
        c.close()


def test_staged_quotient_and_openings(gpu_ctx, oracle):
    """p2g_quotient / p2g_open on their own: fed with the commitments and challenges of a real proof they return
    that proof's quotient chunks, quotient cap and openings"""
    from plonky2_aes_b200.host.polynomial_batch import PolynomialBatch
    from plonky2_aes_b200.host.proof import Proof
    lib = gpu_ctx.lib
    for data, wires, pi in (circuits.aes_gcm(13, True)[:2] + (None,), circuits.public_input_circuit()):
        data.load(gpu_ctx)
        gpu_ctx.check(lib.p2g_set_timing(gpu_ctx.handle, 2))
        proof = data.prove_wires(wires, pi)
        tr = ffi.Transcript()
        gpu_ctx.check(lib.p2g_last_transcript(gpu_ctx.handle, C.byref(tr)))
        d = data.descriptor()
        nch = d.num_challenges
        nlp = 0 if d.num_luts == 0 else -(-(d.num_routed_wires // 2) // (d.quotient_degree_factor - 1)) + 1
        zs = np.empty((nch * (1 + d.num_partial_products + nlp), data.n), dtype=np.uint64)
        qc = np.empty((nch * d.quotient_degree_factor, data.n), dtype=np.uint64)
        gpu_ctx.check(lib.p2g_last_zs_values(gpu_ctx.handle, zs.ctypes.data))
        gpu_ctx.check(lib.p2g_last_quotient_chunks(gpu_ctx.handle, qc.ctypes.data))
        gpu_ctx.check(lib.p2g_set_timing(gpu_ctx.handle, 0))
        wb = PolynomialBatch.from_values(gpu_ctx, wires)
        zb = PolynomialBatch.from_values(gpu_ctx, zs)
        pr = Proof(proof, d)
        assert np.array_equal(wb.cap.ravel(), pr["wires_cap"]) and np.array_equal(zb.cap.ravel(), pr["plonk_zs_partial_products_cap"])
        u64 = lambda xs, k: np.array(list(xs)[:k], dtype=np.uint64)
        betas, gammas, alphas, deltas = u64(tr.betas, nch), u64(tr.gammas, nch), u64(tr.alphas, nch), u64(tr.deltas, 4 * nch)
        h = C.c_void_p()
        cap = np.empty((16, 4), dtype=np.uint64)
        pin = np.ascontiguousarray(pi, dtype=np.uint64) if pi is not None else None
        gpu_ctx.check(lib.p2g_quotient(gpu_ctx.handle, data._gpu_circuit, wb.handle, zb.handle, pin.ctypes.data if pin is not None else None,
                                       betas.ctypes.data, gammas.ctypes.data, deltas.ctypes.data if d.num_luts else None, alphas.ctypes.data,
                                       C.byref(h), cap.ctypes.data))
        assert np.array_equal(cap.ravel(), pr["quotient_polys_cap"])
        qb = PolynomialBatch(gpu_ctx, h, qc.shape[0], data.degree_bits, 3, 4, cap)
        assert np.array_equal(qb.coeffs(), qc)
        # openings at zeta of the wires and quotient batches
        zeta = np.array(list(tr.zeta), dtype=np.uint64)
        out = np.empty((wb.ncols + qb.ncols, 2), dtype=np.uint64)
        gpu_ctx.check(lib.p2g_open(gpu_ctx.handle, (C.c_void_p * 2)(wb.handle, qb.handle), 2, zeta.ctypes.data, out.ctypes.data))
        assert np.array_equal(out[:wb.ncols].ravel(), pr["openings.wires"])
        assert np.array_equal(out[wb.ncols:].ravel(), pr["openings.quotient_polys"])
        # p2g_fri_prove: the transcript replayed on the host (Challenger of iop/challenger.rs over the oracle's
        # permutation) up to the openings, then the FRI part of the proof from the staged entry point
        ch = _Challenger(oracle)
        pi_hash = oracle.hash_no_pad(pin) if pin is not None and len(pin) else np.zeros(4, dtype=np.uint64)
        ch.observe(data.circuit_digest); ch.observe(pi_hash); ch.observe(pr["wires_cap"])
        b_ = [ch.get() for _ in range(nch)]; g_ = [ch.get() for _ in range(nch)]
        assert b_ == list(betas) and g_ == list(gammas)
        if d.num_luts:
            extra = [ch.get() for _ in range(2 * nch)]
            assert b_ + g_ + extra == list(deltas)
        ch.observe(pr["plonk_zs_partial_products_cap"])
        assert [ch.get() for _ in range(nch)] == list(alphas)
        ch.observe(pr["quotient_polys_cap"])
        assert [ch.get(), ch.get()] == list(zeta)
        for name in ("constants", "plonk_sigmas", "wires", "plonk_zs", "partial_products", "quotient_polys", "lookup_zs",
                     "plonk_zs_next", "lookup_zs_next"):
            ch.observe(pr["openings." + name])
        io = ch.export()
        fw = lib.p2g_fri_proof_words(data._gpu_circuit)
        fri = np.empty(fw, dtype=np.uint64)
        got = C.c_size_t()
        gpu_ctx.check(lib.p2g_fri_prove(gpu_ctx.handle, data._gpu_circuit, wb.handle, zb.handle, qb.handle, zeta.ctypes.data,
                                        io.ctypes.data, fri.ctypes.data, fw, C.byref(got)))
        npi = 0 if pi is None else len(pi)
        tail = proof[len(proof) - npi - fw:len(proof) - npi]
        assert got.value == fw and np.array_equal(fri, tail), "FRI part of the proof"
        assert int(io[28]) <= 7 and int(io[29]) <= 8
        # refused: a non-canonical transcript word, a batch of the wrong shape
        bad = io.copy(); bad[3] = P
        assert lib.p2g_fri_prove(gpu_ctx.handle, data._gpu_circuit, wb.handle, zb.handle, qb.handle, zeta.ctypes.data,
                                 bad.ctypes.data, fri.ctypes.data, fw, None) == -2
        assert lib.p2g_fri_prove(gpu_ctx.handle, data._gpu_circuit, zb.handle, wb.handle, qb.handle, zeta.ctypes.data,
                                 io.ctypes.data, fri.ctypes.data, fw, None) == -2
        wb.free(); zb.free(); qb.free()


class _Challenger:
    """Challenger<F, PoseidonHash> (iop/challenger.rs): duplex sponge of rate 8; get() pops the output buffer from its end"""

    def __init__(self, oracle):
        self.o, self.state, self.inb, self.outb = oracle, np.zeros(12, dtype=np.uint64), [], []

    def _duplex(self):
        for i, v in enumerate(self.inb):
            self.state[i] = v
        self.inb = []
        self.state = self.o.poseidon(self.state)
        self.outb = [int(v) for v in self.state[:8]]

    def observe(self, xs):
        for x in np.asarray(xs, dtype=np.uint64).ravel():
            self.outb = []
            self.inb.append(int(x))
            if len(self.inb) == 8:
                self._duplex()

    def get(self):
        if self.inb or not self.outb:
            self._duplex()
        return self.outb.pop()

    def export(self):
        io = np.zeros(30, dtype=np.uint64)
        io[:12] = self.state
        io[12:12 + len(self.inb)] = self.inb
        io[20:20 + len(self.outb)] = self.outb
        io[28], io[29] = len(self.inb), len(self.outb)
        return io
