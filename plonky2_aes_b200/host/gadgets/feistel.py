"""Additive Feistel network over Goldilocks, native and in-circuit — Python mirror of
/root/reference/feistel/src/lib.rs:15-75 (`feistel_cipher`, `feistel_inv_cipher`) and
/root/reference/feistel/src/circuit.rs:21-71 (`CircuitBuilderFeistel::feistel_cipher`)."""
from .poseidon_native import P


def feistel_cipher(state, key_schedule, f):
    half = len(state) // 2
    state = list(state)
    for k in key_schedule:
        l, r = state[:half], state[half:]
        off = f(list(r) + list(k))
        state = r + [(a + b) % P for a, b in zip(l, off)]
    return state


def feistel_inv_cipher(state, reverse_key_schedule, f):
    half = len(state) // 2
    state = list(state)
    for k in reverse_key_schedule:
        l, r = state[:half], state[half:]
        off = f(list(l) + list(k))
        state = [(a - b) % P for a, b in zip(r, off)] + l
    return state


def add_feistel_state_target(builder, state_len):
    return builder.add_virtual_targets(state_len)


def feistel_cipher_target(builder, state, key_schedule, f):
    """f(builder, targets) -> STATE_HALF_LEN targets (e.g. hash_n_to_hash_no_pad)"""
    half = len(state) // 2
    state = list(state)
    for k in key_schedule:
        l, r = state[:half], state[half:]
        off = f(builder, list(r) + list(k))
        state = r + [builder.add(l[i], off[i]) for i in range(half)]
    return state
