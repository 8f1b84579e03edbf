"""GPU parity: PolynomialBatch::from_values/from_coeffs (iNTT, coset LDE in bit-reversed leaf
order, Poseidon Merkle tree, cap, paths) — CUDA path through the C ABI vs the CPU oracle,
bit-exact."""
import numpy as np
import pytest

from plonky2_aes_b200.host.polynomial_batch import PolynomialBatch

P = 0xFFFFFFFF00000001
pytestmark = pytest.mark.gpu


def _cols(seed, ncols, log_n, edge=False):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, P, size=(ncols, 1 << log_n), dtype=np.uint64)
    if edge:
        a[0, :] = P - 1
        a[-1, :] = 0
        a[ncols // 2, ::2] = P - 1
    return a


def _compare(gb, ob, check_levels=True):
    assert np.array_equal(gb.coeffs(), ob.coeffs())
    lde = gb.lde_values()                       # [ncols][N]
    assert np.array_equal(lde.T, ob.leaves())    # oracle: row-major [N][ncols]
    L = gb.path_len
    assert L == ob.tree.path_len
    if check_levels:
        for k in range(L + 1):
            assert np.array_equal(gb.digests(k), ob.tree.level(k)), f"level {k}"
    assert np.array_equal(gb.cap, ob.tree.cap)


@pytest.mark.parametrize("ncols,log_n", [(1, 1), (3, 2), (4, 4), (5, 5), (9, 6), (16, 7), (34, 9), (2, 10), (7, 12), (3, 13), (3, 14), (2, 16), (5, 17)])
def test_from_values_small(gpu_ctx, oracle, ncols, log_n):
    cols = _cols(10 + log_n, ncols, log_n, edge=True)
    gb = PolynomialBatch.from_values(gpu_ctx, cols, 3, min(4, log_n + 3))
    ob = oracle.batch(cols, True, 3, min(4, log_n + 3))
    _compare(gb, ob)
    gb.free(); ob.free()


@pytest.mark.parametrize("ncols,log_n", [(135, 13), (20, 14), (12, 15), (3, 16)])
def test_from_values_reference_sizes(gpu_ctx, oracle, ncols, log_n):
    cols = _cols(100 + log_n, ncols, log_n)
    gb = PolynomialBatch.from_values(gpu_ctx, cols)
    ob = oracle.batch(cols, True)
    _compare(gb, ob)
    # MerkleTree::get + prove
    for j in (0, 12345, gb.lde_size - 1):
        row, sib = gb.get_and_prove(j)
        assert np.array_equal(row, ob.leaves()[j])
        assert np.array_equal(sib, ob.tree.prove(j))
    gb.free(); ob.free()


@pytest.mark.parametrize("rate_bits,cap_height", [(1, 0), (2, 3), (3, 4), (4, 2)])
def test_from_coeffs_rates_and_caps(gpu_ctx, oracle, rate_bits, cap_height):
    cols = _cols(7, 16, 8)
    gb = PolynomialBatch.from_coeffs(gpu_ctx, cols, rate_bits, cap_height)
    ob = oracle.batch(cols, False, rate_bits, cap_height)
    _compare(gb, ob)
    gb.free(); ob.free()


def test_full_size_wires_commitment_properties(gpu_ctx, oracle):
    """BASELINE config 2 shape: 135 columns, n = 2^15.  Size-independent checks: LDE restricted
    to the subgroup coset reproduces Horner evaluation; linearity of the LDE; cap equals the
    oracle's (the oracle finishes this size in seconds)."""
    cols = _cols(2026, 135, 15)
    gb = PolynomialBatch.from_values(gpu_ctx, cols)
    ob = oracle.batch(cols, True)
    assert np.array_equal(gb.cap, ob.tree.cap)
    assert np.array_equal(gb.digests(0)[::4097], ob.tree.level(0)[::4097])
    lde = gb.lde_values()
    assert np.array_equal(lde[:, ::1023].T, ob.leaves()[::1023])
    # linearity: commit(a) + commit(b) == commit(a+b) pointwise on the LDE
    a, b = cols[:4], cols[4:8]
    s = ((a.astype(object) + b.astype(object)) % P).astype(np.uint64)
    gs = PolynomialBatch.from_values(gpu_ctx, s)
    lhs = ((lde[:4].astype(object) + lde[4:8].astype(object)) % P).astype(np.uint64)
    assert np.array_equal(gs.lde_values(), lhs)
    gs.free(); gb.free(); ob.free()


def test_hash_and_merkle_helpers(gpu_ctx, oracle):
    rng = np.random.default_rng(5)
    for ll in (1, 4, 5, 8, 9, 16, 32, 135):
        rows = rng.integers(0, P, size=(64, ll), dtype=np.uint64)
        got = gpu_ctx.hash_no_pad_many(rows)
        for i in (0, 63):
            assert list(got[i]) == list(oracle.hash_no_pad(rows[i]))
        cap, dig = gpu_ctx.merkle_cap(rows, 4)
        t = oracle.merkle(rows, 4)
        assert np.array_equal(cap, t.cap)
        assert np.array_equal(dig, t.level(0))
        t.free()
    # poseidon KAT through the device: hash of 8 zeros = permutation(0)[0..4]
    z = gpu_ctx.hash_no_pad_many(np.zeros((1, 8), dtype=np.uint64))
    assert [hex(int(v)) for v in z[0]] == ['0x3c18a9786cb0b359', '0xc4055e3364a246c3', '0x7953db0ab48808f4', '0xc71603f33a1144ca']


def test_poseidon_extreme_words(gpu_ctx, oracle):
    """The FP64 linear layers are exact only while their accumulators stay inside [2^52, 2^53): states made of the words
    that maximise / minimise the 32-bit halves (0, p - 1, 2^32 - 1, 2^64 - 2^32, 2^32, 1) in the patterns that
    drive X+ / X- and the rank-one term of the merged partial rounds to their extremes, plus 2000 random 8-word rows,
    through the thread-per-permutation kernel and the 12-lane kernel (merkle_cap of a narrow tree)."""
    ext = [0, P - 1, 0xFFFFFFFF, 0xFFFFFFFF00000000, 1 << 32, 1, 0xFFFFFFFEFFFFFFFF, 0x7FFFFFFF80000000]
    rows = []
    for a in ext:
        for b in ext:
            rows.append([a] * 4 + [b] * 4)                  # word j against word j + 4
            rows.append([a, b] * 4)                         # alternating
            rows.append([a] + [b] * 7)                      # word 0 (the partial-round S-box) against the rest
    rows = np.array(rows, dtype=np.uint64)
    rng = np.random.default_rng(77)
    rows = np.concatenate([rows, rng.integers(0, P, size=(2000, 8), dtype=np.uint64)])
    got = gpu_ctx.hash_no_pad_many(rows)
    for i in range(len(rows)):
        assert list(got[i]) == list(oracle.hash_no_pad(rows[i])), i
    # the same words as 16-word leaves of a 64-leaf tree: leaf sponge (2 permutations) + 12-lane levels
    leaves = np.concatenate([rows[:64], rows[64:128]], axis=1)
    cap, dig = gpu_ctx.merkle_cap(leaves, 2)
    t = oracle.merkle(leaves, 2)
    assert np.array_equal(cap, t.cap) and np.array_equal(dig, t.level(0))
    t.free()


@pytest.mark.parametrize("ncols,log_n", [(20, 10), (3, 14), (2, 17)])
def test_coset_sharded_commit_matches_full(gpu_ctx, oracle, ncols, log_n):
    """multi-GPU coset split of one commitment, emulated on one GPU: every shard's LDE block and cap
    entries equal the corresponding slice of the full commitment (and of the oracle's).  log_n = 14
    takes the 2^13-point / two-blocks-per-SM NTT configuration, log_n = 17 the pre-folded one (outer kernel on a
    subset of the cosets)."""
    import ctypes as C
    import torch
    from plonky2_aes_b200.host.sharding import sharded_commit
    cols = _cols(77, ncols, log_n)
    ob = oracle.batch(cols, True)
    dev = torch.from_numpy(cols.view(np.int64)).cuda()
    N = 8 << log_n
    for world in (1, 2, 4, 8):
        caps = []
        for rank in range(world):
            part, h = sharded_commit(gpu_ctx, dev, ncols, log_n, rank, world)
            caps.append(part)
            per = 8 // world
            lde = np.empty((ncols, per << log_n), dtype=np.uint64)
            gpu_ctx.check(gpu_ctx.lib.p2g_batch_get_lde(gpu_ctx.handle, h, lde.ctypes.data))
            assert np.array_equal(lde.T, ob.leaves()[rank * (N // world):(rank + 1) * (N // world)])
            gpu_ctx.check(gpu_ctx.lib.p2g_batch_free(gpu_ctx.handle, h))
        assert np.array_equal(np.concatenate(caps), ob.tree.cap), world
    ob.free()


@pytest.mark.parametrize("log_leaves", [1, 2, 3, 4])
def test_merkle_fewer_leaves_than_a_warp(gpu_ctx, oracle, log_leaves):
    """2..16 leaves with cap heights below the leaf level: the leaf kernel's block is padded to a full
    warp and must not fold (or write) nodes that do not exist."""
    rng = np.random.default_rng(40 + log_leaves)
    for leaf_len in (3, 9):
        rows = rng.integers(0, P, size=(1 << log_leaves, leaf_len), dtype=np.uint64)
        for cap_height in range(0, log_leaves + 1):
            cap, dig = gpu_ctx.merkle_cap(rows, cap_height)
            t = oracle.merkle(rows, cap_height)
            assert np.array_equal(cap, t.cap), (log_leaves, leaf_len, cap_height)
            assert np.array_equal(dig, t.level(0))
            t.free()
    cols = _cols(log_leaves, 5, 1)
    for cap_height in (0, 1, 2, 3):
        gb = PolynomialBatch.from_values(gpu_ctx, cols, 3, cap_height)
        ob = oracle.batch(cols, True, 3, cap_height)
        _compare(gb, ob)
        gb.free(); ob.free()


def test_unsupported_sizes_are_refused(gpu_ctx):
    """transforms above 2^20 points are refused, not attempted"""
    import ctypes as C
    h = C.c_void_p()
    dummy = np.zeros(8, dtype=np.uint64)
    rc = gpu_ctx.lib.p2g_commit_from_values(gpu_ctx.handle, dummy.ctypes.data, 1, 21, 3, 4, C.byref(h), None)
    assert rc == -2


@pytest.mark.parametrize("ncols,log_n", [(3, 18), (1, 20)])
def test_from_values_large_degrees(gpu_ctx, oracle, ncols, log_n):
    """n >= 2^17 runs the pre-folded NTT (outer R-point stages in a kernel of their own, R = n / 2^14 up to 64)"""
    cols = _cols(300 + log_n, ncols, log_n, edge=ncols > 1)
    gb = PolynomialBatch.from_values(gpu_ctx, cols)
    ob = oracle.batch(cols, True)
    _compare(gb, ob, check_levels=False)
    for j in (0, 54321, gb.lde_size - 1):
        row, sib = gb.get_and_prove(j)
        assert np.array_equal(row, ob.leaves()[j])
        assert np.array_equal(sib, ob.tree.prove(j))
    gb.free(); ob.free()


_PREFOLD_SCRIPT = """
import sys
import numpy as np
sys.path.insert(0, %r)
from tests import oracle_lib
from tests.test_gpu_commit import _cols, _compare
from plonky2_aes_b200.host.polynomial_batch import Context, PolynomialBatch
oracle = oracle_lib.load()
ctx = Context(0)
for ncols, log_n, rate, from_values in %r:
    cols = _cols(500 + log_n, ncols, log_n, edge=True)
    gb = (PolynomialBatch.from_values if from_values else PolynomialBatch.from_coeffs)(ctx, cols, rate, min(4, log_n))
    ob = oracle.batch(cols, from_values, rate, min(4, log_n))
    _compare(gb, ob)
    gb.free(); ob.free()
ctx.close()
print("prefold ok")
"""


@pytest.mark.parametrize("log_m,cases", [(5, [(3, 6, 3, True), (2, 8, 3, True), (4, 12, 1, False)]),
                                         (9, [(3, 10, 3, True), (5, 13, 3, True), (2, 16, 2, True)])])
def test_prefolded_ntt_forced_at_small_sizes(log_m, cases):
    """The pre-folded shape (R-point outer kernel + size-M kernel with a one-word-per-point table) forced through
    P2G_NTT_PREFOLD / P2G_NTT_LOG_M at sizes the oracle checks in full: R = 2 ... 128, inverse and coset forms.
    Runs in a subprocess because NTT plans are cached per device for the life of the process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, P2G_NTT_PREFOLD="1", P2G_NTT_LOG_M=str(log_m))
    r = subprocess.run([sys.executable, "-c", _PREFOLD_SCRIPT % (root, cases)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "prefold ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_contexts_on_two_devices_in_one_process(oracle):
    """One process driving several GPUs (the Rust side's model): the > 48 KB shared-memory opt-in of the
    NTT kernel is per device.  Needs 2 GPUs; with one GPU two contexts on it are checked instead."""
    import torch
    from plonky2_aes_b200.host.polynomial_batch import Context
    ndev = torch.cuda.device_count()
    devs = [0, 1] if ndev >= 2 else [0, 0]
    cols = _cols(9, 3, 13)                   # log_m = 13: 69 KB of dynamic shared memory
    ob = oracle.batch(cols, True)
    ctxs = [Context(d) for d in devs]
    for c in ctxs:
        gb = PolynomialBatch.from_values(c, cols)
        assert np.array_equal(gb.cap, ob.tree.cap)
        gb.free()
    for c in ctxs:
        c.close()
    ob.free()
