"""CPU tests that pin the oracle: recalled upstream Poseidon vectors (the only known-answer
tests that exist for this path — SURVEY.md §8c: the reference's own tests pin no value on the
prove() path), the round-constant generator, and algebraic identities of the FFT/LDE/Merkle
restatement."""
import json
import os

import numpy as np

P = 0xFFFFFFFF00000001
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_poseidon_round_constants_regenerate():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLD), "..", "tools"))
    from gen_poseidon_constants import generate, RECALLED_HEAD
    c = generate()
    assert c[:len(RECALLED_HEAD)] == RECALLED_HEAD
    with open(os.path.join(GOLD, "poseidon_kat.json")) as f:
        kat = json.load(f)
    assert [hex(v) for v in c] == kat["round_constants"]


def test_poseidon_kat_upstream(oracle):
    # plonky2/src/hash/poseidon_goldilocks.rs `test_vectors` (recalled): zero state, 0..11
    with open(os.path.join(GOLD, "poseidon_kat.json")) as f:
        kat = json.load(f)
    for vec in kat["permutation"]:
        out = oracle.poseidon([int(x, 16) for x in vec["input"]])
        assert [hex(int(v)) for v in out] == vec["output"]
        assert list(oracle.poseidon_naive([int(x, 16) for x in vec["input"]])) == list(out)
    rng = np.random.default_rng(11)
    for _ in range(50):
        s = rng.integers(0, P, size=12, dtype=np.uint64)
        assert list(oracle.poseidon(s)) == list(oracle.poseidon_naive(s))


def test_hash_sponge_semantics(oracle):
    rng = np.random.default_rng(1)
    x = rng.integers(0, P, size=20, dtype=np.uint64)
    # overwrite mode: 20 inputs = chunks 8, 8, 4; squeeze 20 outputs = 8, 8, 4
    s = np.zeros(12, dtype=np.uint64)
    for off in (0, 8, 16):
        chunk = x[off:off + 8]
        s[:len(chunk)] = chunk
        s = oracle.poseidon(s)
    exp = list(s[:8]); s = oracle.poseidon(s); exp += list(s[:8]); s = oracle.poseidon(s); exp += list(s[:4])
    assert list(oracle.hash_n_to_m_no_pad(x, 20)) == exp
    # two_to_one == permutation of [l, r, 0,0,0,0]
    l, r = x[:4], x[4:8]
    st = np.concatenate([l, r, np.zeros(4, dtype=np.uint64)])
    assert list(oracle.two_to_one(l, r)) == list(oracle.poseidon(st)[:4])


def _horner(coeffs, x):
    acc = 0
    for c in reversed([int(v) for v in coeffs]):
        acc = (acc * x + c) % P
    return acc


def test_fft_roundtrip_and_definition(oracle):
    rng = np.random.default_rng(2)
    for log_n in (0, 1, 3, 8):
        n = 1 << log_n
        c = rng.integers(0, P, size=n, dtype=np.uint64)
        v = oracle.fft(c)
        assert list(oracle.fft(v, inverse=True)) == list(c)
        w = pow(1753635133440165772, 1 << (32 - log_n), P)
        for i in (0, n // 2, n - 1):
            assert int(v[i]) == _horner(c, pow(w, i, P))
        cv = oracle.coset_fft(c, 7)
        assert int(cv[n - 1]) == _horner(c, 7 * pow(w, n - 1, P) % P)
        assert list(oracle.coset_fft(cv, 7, inverse=True)) == list(c)


def test_batch_leaves_are_coset_evaluations(oracle):
    rng = np.random.default_rng(3)
    log_n, ncols = 5, 7
    n, N = 1 << log_n, 1 << (log_n + 3)
    vals = rng.integers(0, P, size=(ncols, n), dtype=np.uint64)
    b = oracle.batch(vals, from_values=True)
    co = b.coeffs()
    g = pow(1753635133440165772, 1 << (32 - log_n), P)
    for c in (0, ncols - 1):
        for i in (0, 1, n - 1):
            assert _horner(co[c], pow(g, i, P)) == int(vals[c, i])
    leaves = b.leaves()
    wN = pow(1753635133440165772, 1 << (32 - log_n - 3), P)
    for j in (0, 1, 5, N - 1):
        i = int(format(j, f"0{log_n + 3}b")[::-1], 2)
        x = 7 * pow(wN, i, P) % P
        for c in (0, 3, ncols - 1):
            assert int(leaves[j, c]) == _horner(co[c], x)
    # Merkle: every path verifies against the cap, and digests follow hash_or_noop / two_to_one
    t = b.tree
    assert list(t.level(0)[5]) == list(oracle.hash_no_pad(leaves[5]))
    assert list(t.level(1)[2]) == list(oracle.two_to_one(t.level(0)[4], t.level(0)[5]))
    cap = t.cap
    assert cap.shape == (16, 4)
    for j in (0, 77, N - 1):
        sib = t.prove(j)
        leaf = np.ascontiguousarray(leaves[j])
        ok = oracle.lib.orc_merkle_verify(leaf.ctypes.data, ncols, j, cap.ctypes.data, 4, sib.ctypes.data, len(sib))
        assert ok == 1
        bad = leaf.copy(); bad[0] ^= 1
        assert oracle.lib.orc_merkle_verify(bad.ctypes.data, ncols, j, cap.ctypes.data, 4, sib.ctypes.data, len(sib)) == 0
    b.free()


def test_hash_or_noop_small_leaves(oracle):
    leaves = np.arange(32, dtype=np.uint64).reshape(16, 2)
    t = oracle.merkle(leaves, 2)
    assert list(t.level(0)[3]) == [6, 7, 0, 0]
    t.free()
