"""Pinning against the real prover.  Every `tests/golden/reference_*.p2gfix` file (written by
tools/dump_reference_fixture.rs on a machine with the Rust reference) is replayed: the dumped circuit
description and wire matrix go to the CPU oracle (and, with -m gpu, through the C ABI to the CUDA prover),
and the proof BYTES must equal the ones plonky2 @ 109d517 produced.  Until such a file exists the replay
tests are skipped -- loudly -- and only the loader itself is exercised on a self-generated fixture."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from plonky2_aes_b200.host.proof import Proof
from tests import circuits, fixture_format, oracle_lib

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "reference_*.p2gfix")))
NO_FIXTURE = ("PARITY UNPINNED: no tests/golden/reference_*.p2gfix present. Run tools/dump_reference_fixture.rs where the Rust "
              "reference builds (see its header) and drop the files into tests/golden/ to pin the oracle and the CUDA prover "
              "to plonky2 @ 109d517.")


def _replay_oracle(orc, fc):
    oc = oracle_lib.OracleCircuit(orc, fc)
    assert np.array_equal(oc.cap, fc.constants_sigmas_cap), "constants_sigmas cap differs from the reference's"
    words = oc.prove(fc.wires, public_inputs=fc.public_inputs)
    assert Proof(words, fc.descriptor()).to_bytes() == fc.proof_bytes, "oracle proof bytes differ from the reference's"
    ref_words = Proof.from_bytes(fc.proof_bytes, fc.descriptor()).words
    assert oc.verify(ref_words) == 0, "restated verifier rejects the reference's own proof"
    oc.free()


def _replay_gpu(ctx, fc):
    from plonky2_aes_b200.host.circuit_builder import CircuitData
    lib, d = ctx.lib, fc.descriptor()
    h = C.c_void_p()
    cap = np.empty_like(fc.constants_sigmas_cap)
    ctx.check(lib.p2g_circuit_load(ctx.handle, C.byref(d), C.byref(h), cap.ctypes.data))
    try:
        assert np.array_equal(cap, fc.constants_sigmas_cap), "GPU constants_sigmas cap differs from the reference's"
        nw = lib.p2g_proof_words(h)
        words, got = np.empty(nw, dtype=np.uint64), C.c_size_t()
        pi = fc.public_inputs if fc.public_inputs.size else None
        ctx.check(lib.p2g_prove(ctx.handle, h, fc.wires.ctypes.data, pi.ctypes.data if pi is not None else None,
                                words.ctypes.data, nw, C.byref(got)))
        assert Proof(words[:got.value], d).to_bytes() == fc.proof_bytes, "GPU proof bytes differ from the reference's"
    finally:
        lib.p2g_circuit_free(ctx.handle, h)
    del CircuitData


@pytest.mark.skipif(bool(FIXTURES), reason="reference fixtures present: the replay tests below run instead")
def test_reference_fixture_absent_is_reported():
    pytest.skip(NO_FIXTURE)


@pytest.mark.parametrize("path", FIXTURES or [None])
def test_oracle_reproduces_reference_proof_bytes(oracle, path):
    if path is None:
        pytest.skip(NO_FIXTURE)
    fc = fixture_format.FixtureCircuit(fixture_format.read(path))
    assert not fc.unsupported, f"gates the backend does not evaluate: {fc.unsupported}"
    _replay_oracle(oracle, fc)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES or [None])
def test_gpu_reproduces_reference_proof_bytes(gpu_ctx, path):
    if path is None:
        pytest.skip(NO_FIXTURE)
    fc = fixture_format.FixtureCircuit(fixture_format.read(path))
    _replay_gpu(gpu_ctx, fc)


def _self_fixture(oracle, tmp_path, with_public_inputs):
    if with_public_inputs:
        data, wires, pi = circuits.public_input_circuit()
    else:
        data, wires, _ = circuits.aes_gcm(13, True)
        pi = np.zeros(0, dtype=np.uint64)
    oracle_lib.set_circuit_digest(oracle, data)
    oc = oracle_lib.OracleCircuit(oracle, data)
    words = oc.prove(wires, public_inputs=pi)
    oc.free()
    path = str(tmp_path / "self.p2gfix")
    fixture_format.write(path, fixture_format.from_circuit_data(data, wires, Proof(words, data.descriptor()).to_bytes(), pi))
    return fixture_format.FixtureCircuit(fixture_format.read(path))


@pytest.mark.parametrize("with_pi", [False, True])
def test_fixture_loader_on_self_generated_file(oracle, tmp_path, with_pi):
    """the replay machinery end to end (file format, descriptor from a file, byte comparison) on a fixture
    this repo wrote itself: proves the loader, NOT parity with the reference"""
    fc = _self_fixture(oracle, tmp_path, with_pi)
    _replay_oracle(oracle, fc)
    bad = fixture_format.FixtureCircuit(fc.fx)
    bad.proof_bytes = bad.proof_bytes[:100] + bytes([bad.proof_bytes[100] ^ 1]) + bad.proof_bytes[101:]
    with pytest.raises(AssertionError):
        _replay_oracle(oracle, bad)


@pytest.mark.gpu
def test_fixture_loader_gpu_on_self_generated_file(gpu_ctx, oracle, tmp_path):
    _replay_gpu(gpu_ctx, _self_fixture(oracle, tmp_path, False))
    _replay_gpu(gpu_ctx, _self_fixture(oracle, tmp_path, True))
