"""ctypes binding of include/p2gpu.h (the same symbols a Rust `extern "C"` block would bind;
see INTEGRATION.md).  Host-side mirror only: every call goes to libp2gpu.so."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

P2G_ERRORS = {-1: "P2G_E_CUDA", -2: "P2G_E_BADARG", -3: "P2G_E_UNSAT", -4: "P2G_E_POW", -5: "P2G_E_NOMEM"}


class P2GError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__(f"{P2G_ERRORS.get(code, code)}: {msg}")


def lib_path():
    # P2G_LIB_PATH: developer override used to A/B kernel variants; the default is the in-tree build
    return os.environ.get("P2G_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "libp2gpu.so")


class Gate(C.Structure):
    _fields_ = [("kind", C.c_int32), ("selector_index", C.c_int32), ("group_start", C.c_int32),
                ("group_end", C.c_int32), ("num_constraints", C.c_int32), ("param0", C.c_int32)]


class CircuitDesc(C.Structure):
    _fields_ = [
        ("degree_bits", C.c_int32),
        ("num_wires", C.c_int32), ("num_routed_wires", C.c_int32), ("num_constants", C.c_int32),
        ("num_challenges", C.c_int32), ("quotient_degree_factor", C.c_int32),
        ("rate_bits", C.c_int32), ("cap_height", C.c_int32), ("pow_bits", C.c_int32), ("num_query_rounds", C.c_int32),
        ("num_reduction_arity_bits", C.c_int32), ("reduction_arity_bits", C.c_int32 * 16),
        ("num_selectors", C.c_int32), ("num_lookup_selectors", C.c_int32),
        ("num_gates", C.c_int32), ("gates", C.POINTER(Gate)),
        ("num_gate_constraints", C.c_int32),
        ("num_partial_products", C.c_int32),
        ("num_luts", C.c_int32),
        ("lut_lens", C.POINTER(C.c_int32)),
        ("lut_data", C.POINTER(C.c_uint16)),
        ("lookup_rows", C.POINTER(C.c_int32)),
        ("num_public_inputs", C.c_int32),
        ("k_is", C.POINTER(C.c_uint64)),
        ("constants_sigmas", C.POINTER(C.c_uint64)),
        ("circuit_digest", C.c_uint64 * 4),
    ]


class Transcript(C.Structure):
    _fields_ = [("betas", C.c_uint64 * 4), ("gammas", C.c_uint64 * 4), ("deltas", C.c_uint64 * 16),
                ("alphas", C.c_uint64 * 4), ("zeta", C.c_uint64 * 2), ("fri_alpha", C.c_uint64 * 2),
                ("fri_betas", C.c_uint64 * 32), ("pow_witness", C.c_uint64), ("query_indices", C.c_uint64 * 64)]


class Timings(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("h2d", "wires_commit", "zs_build", "zs_commit", "quotient", "quotient_commit",
                                         "openings", "fri_combine", "fri_commit", "pow", "queries", "total")]


# every symbol include/p2gpu.h declares, with its signature
_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p
SIGNATURES = {
    "p2g_version": (C.c_int32, []),
    "p2g_ctx_create": (C.c_int32, [C.c_int32, C.POINTER(_vp)]),
    "p2g_ctx_destroy": (None, [_vp]),
    "p2g_last_error": (C.c_char_p, [_vp]),
    "p2g_ctx_sync": (C.c_int32, [_vp]),
    "p2g_ctx_stream": (_vp, [_vp]),
    "p2g_commit_from_values": (C.c_int32, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_vp), _vp]),
    "p2g_commit_from_coeffs": (C.c_int32, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_vp), _vp]),
    "p2g_commit_from_values_dev": (C.c_int32, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_vp), _vp]),
    "p2g_commit_from_coeffs_dev": (C.c_int32, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_vp), _vp]),
    "p2g_commit_blocks_from_values_dev": (C.c_int32, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_vp), _vp]),
    "p2g_batch_free": (C.c_int32, [_vp, _vp]),
    "p2g_batch_get_coeffs": (C.c_int32, [_vp, _vp, _vp]),
    "p2g_batch_get_lde": (C.c_int32, [_vp, _vp, _vp]),
    "p2g_batch_get_level": (C.c_int32, [_vp, _vp, C.c_uint32, _vp]),
    "p2g_batch_open_leaf": (C.c_int32, [_vp, _vp, C.c_uint64, _vp, _vp]),
    "p2g_merkle_cap": (C.c_int32, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp]),
    "p2g_hash_no_pad_many": (C.c_int32, [_vp, _vp, C.c_uint32, C.c_uint32, _vp]),
    "p2g_circuit_load": (C.c_int32, [_vp, C.POINTER(CircuitDesc), C.POINTER(_vp), _vp]),
    "p2g_circuit_free": (C.c_int32, [_vp, _vp]),
    "p2g_proof_words": (C.c_size_t, [_vp]),
    "p2g_wprog_load": (C.c_int32, [_vp, _vp, _vp, C.c_uint32, C.POINTER(_vp)]),
    "p2g_wprog_free": (C.c_int32, [_vp, _vp]),
    "p2g_wprog_ext_slots": (C.c_uint32, [_vp]),
    "p2g_wprog_levels": (C.c_uint32, [_vp]),
    "p2g_wprog_generate": (C.c_int32, [_vp, _vp, _vp, C.c_uint32, _vp]),
    "p2g_wprog_generate_dev": (C.c_int32, [_vp, _vp, _vp, C.c_uint32, _vp, _vp]),
    "p2g_prove_slots_dev": (C.c_int32, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "p2g_prove_inputs": (C.c_int32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "p2g_proof_bytes_len": (C.c_size_t, [_vp]),
    "p2g_proof_to_bytes": (C.c_int32, [_vp, _vp, C.c_size_t, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "p2g_proof_from_bytes": (C.c_int32, [_vp, _vp, C.c_size_t, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "p2g_prove": (C.c_int32, [_vp, _vp, _vp, _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "p2g_shard_buffer_bytes": (C.c_size_t, [_vp, C.c_uint32]),
    "p2g_prove_sharded": (C.c_int32, [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, _vp, _vp, C.c_size_t, _vp, _vp, _vp, C.c_size_t,
                                      C.POINTER(C.c_size_t)]),
    "p2g_debug_canary": (C.c_int32, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "p2g_quotient": (C.c_int32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp), _vp]),
    "p2g_open": (C.c_int32, [_vp, _vp, C.c_uint32, _vp, _vp]),
    "p2g_fri_proof_words": (C.c_size_t, [_vp]),
    "p2g_fri_prove": (C.c_int32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "p2g_prove_batch": (C.c_int32, [_vp, _vp, C.c_uint32, _vp, _vp, C.c_uint32, _vp, C.c_size_t, _vp]),
    "p2g_prove_dev": (C.c_int32, [_vp, _vp, _vp, _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "p2g_last_transcript": (C.c_int32, [_vp, C.POINTER(Transcript)]),
    "p2g_last_zs_values": (C.c_int32, [_vp, _vp]),
    "p2g_last_quotient_chunks": (C.c_int32, [_vp, _vp]),
    "p2g_last_timings": (C.c_int32, [_vp, C.POINTER(Timings)]),
    "p2g_set_timing": (C.c_int32, [_vp, C.c_int32]),
    "p2g_last_commit_timings": (C.c_int32, [_vp, C.POINTER(C.c_float * 3)]),
    "p2g_launch_count": (C.c_uint64, []),
    "p2g_pow_grind": (C.c_int32, [_vp, _vp, C.c_uint32, C.c_uint32, _vp]),
    "p2g_fri_fold": (C.c_int32, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint64, _vp, _vp]),
    "p2g_poseidon_peak": (C.c_int32, [_vp, C.c_uint32, C.POINTER(C.c_double)]),
    "p2g_field_ops": (C.c_int32, [_vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "p2g_wmap_load": (C.c_int32, [_vp, _vp, _vp, C.c_uint32, _vp, _vp, C.c_uint32, C.POINTER(_vp)]),
    "p2g_wmap_free": (C.c_int32, [_vp, _vp]),
    "p2g_prove_slots": (C.c_int32, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "p2g_wmap_fill": (C.c_int32, [_vp, _vp, _vp, _vp, _vp]),
}


def load_library():
    """Load libp2gpu.so (fails loudly when it has not been built: there is no fallback)."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise P2GError(-1, f"{path} not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB
