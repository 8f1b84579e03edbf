"""Poseidon-Goldilocks permutation and sponge on Python ints (witness source for the Feistel /
Poseidon gadgets; mirrors what the reference calls through plonky2:
`hash_n_to_hash_no_pad::<F, PoseidonPermutation<_>>`, /root/reference/feistel/src/lib.rs:106)."""
import os
import re

P = 0xFFFFFFFF00000001
_CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]


def _load_rc():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "csrc", "poseidon_rc.inc")
    return [int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{16})ULL", open(path).read())]


RC = _load_rc()
assert len(RC) == 360


def permute(state):
    s = [int(x) % P for x in state]
    r = 0
    for n, full in ((4, True), (22, False), (4, True)):
        for _ in range(n):
            s = [(s[i] + RC[12 * r + i]) % P for i in range(12)]
            if full:
                s = [pow(x, 7, P) for x in s]
            else:
                s[0] = pow(s[0], 7, P)
            s = [(sum(s[(i + q) % 12] * _CIRC[i] for i in range(12)) + (8 * s[0] if q == 0 else 0)) % P for q in range(12)]
            r += 1
    return s


def hash_n_to_m_no_pad(inputs, m):
    s = [0] * 12
    inputs = [int(x) % P for x in inputs]
    for k in range(0, len(inputs), 8):
        chunk = inputs[k:k + 8]
        s[:len(chunk)] = chunk
        s = permute(s)
    out = []
    while True:
        for x in s[:8]:
            out.append(x)
            if len(out) == m:
                return out
        s = permute(s)


def hash_n_to_hash_no_pad(inputs):
    return hash_n_to_m_no_pad(inputs, 4)
