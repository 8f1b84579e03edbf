"""AES-GCM circuit (GCTR + GHASH + tag) — Python mirror of `AesGcmTarget<NK,NB,NR,L,TAG>`
(/root/reference/aes-gcm/src/circuit_gcm.rs:24-208) and its helper gadgets (:212-425)."""
from .aes import (AESStateOps, byte_xor, byte_xor_lut, flatten, from_flat, gf_2_8_mul_lut, sbox_lut,
                  state_mix_matrix)
from .native import TAG_LEN
from ..circuit_builder import BoolTarget, P


def u8_unit_right_shift_lut(builder):
    return builder.add_lookup_table_from_pairs([(x, x >> 1) for x in range(256)])


def u8_bitref_lut(builder):
    return builder.add_lookup_table_from_pairs([((x << 3) + i, (x >> i) & 1) for x in range(256) for i in range(8)])


def u8_bitref(builder, lut_idx, x, i):
    idx = builder.mul_const_add(8, x, i)
    return BoolTarget(builder.add_lookup_from_index(idx, lut_idx))


def xor_blocks(builder, xor_lut_idx, b1, b2):
    return [byte_xor(builder, xor_lut_idx, b1[i], b2[i]) for i in range(16)]


def inc32_target(builder, block):
    r = list(block)
    zero = builder.zero()
    u8_max = builder.constant(255)
    carry = builder.one()
    for byte_index in (15, 14, 13, 12):
        a = block[byte_index]
        s = builder.add(a, carry)
        a_is_max = builder.is_equal(a, u8_max)
        carry_out = builder.mul(carry, a_is_max.target)
        r[byte_index] = builder.select(a_is_max, zero, s)
        carry = carry_out
    return r


def right_shift_one_target(builder, shift_lut_idx, v):
    r = list(v)
    carry = builder.zero()
    for i in range(16):
        current = v[i]
        shifted = builder.add_lookup_from_index(current, shift_lut_idx)
        next_carry = builder.mul_const_add(P - 2, shifted, current)
        r[i] = builder.mul_const_add(1 << 7, carry, shifted)
        carry = next_carry
    return r


def gf_2_128_mul_target(builder, xor_lut_idx, shift_lut_idx, bitref_lut_idx, x, y):
    zero = builder.zero()
    r_first = builder.constant(225)
    z = [zero] * 16
    v = list(y)
    for i in range(128):
        byte_index, bit_index = i // 8, 7 - (i % 8)
        xi = u8_bitref(builder, bitref_lut_idx, x[byte_index], builder.constant(bit_index))
        for b in range(16):
            z_xor_v = byte_xor(builder, xor_lut_idx, z[b], v[b])
            z[b] = builder.select(xi, z_xor_v, z[b])
        lsb = u8_bitref(builder, bitref_lut_idx, v[15], zero)
        v = right_shift_one_target(builder, shift_lut_idx, v)
        v_xor_r = byte_xor(builder, xor_lut_idx, v[0], r_first)
        v[0] = builder.select(lsb, v_xor_r, v[0])
    return z


def ghash_target(builder, xor_lut_idx, shift_lut_idx, bitref_lut_idx, h, x):
    assert len(x) % 16 == 0
    y = [builder.zero()] * 16
    for i in range(len(x) // 16):
        y_xi = xor_blocks(builder, xor_lut_idx, y, x[16 * i:16 * i + 16])
        y = gf_2_128_mul_target(builder, xor_lut_idx, shift_lut_idx, bitref_lut_idx, y_xi, h)
    return y


def gctr_target(builder, ops, nr, luts, mix_matrix, key, icb, x):
    xor_lut, gf_lut, sb_lut = luts
    L = len(x)
    y = list(x)
    cb = list(icb)
    zero = builder.zero()
    for i in range(0, L, 16):
        if i > 0:
            cb = inc32_target(builder, cb)
        raw = x[i:i + 16]
        x_i = raw + [zero] * (16 - len(raw))
        ks = flatten(ops.encrypt_block(nr, xor_lut, gf_lut, sb_lut, mix_matrix, from_flat(cb), key))
        if len(raw) == 16:
            y_i, nb = xor_blocks(builder, xor_lut, x_i, ks), 16
        else:
            m = ks[:L % 16] + [zero] * (16 - L % 16)
            y_i, nb = xor_blocks(builder, xor_lut, x_i, m), L % 16
        y[i:i + nb] = y_i[:nb]
    return y


class AesGcmTarget:
    """AesGcmTarget::<NK, 4, NR, L, TAG>::build(builder) / set_targets(pw, key, nonce, pt, ct, tag)."""

    def __init__(self, builder, nk=4, nr=10, L=16, tag=True):
        self.nk, self.nr, self.L, self.with_tag = nk, nr, L, tag
        ops = AESStateOps(builder)
        sb, xo, gf = sbox_lut(builder), byte_xor_lut(builder), gf_2_8_mul_lut(builder)
        self.key = [ops.add_virtual_byte_target(sb) for _ in range(nk * 4)]
        self.nonce = [ops.add_virtual_byte_target(sb) for _ in range(12)]
        self.pt = [ops.add_virtual_byte_target(sb) for _ in range(L)]
        self.tag = [ops.add_virtual_byte_target(sb) for _ in range(TAG_LEN // 8)]
        mix = state_mix_matrix(builder)
        w = ops.key_expansion(nk, nr, xo, sb, self.key)
        h = flatten(ops.encrypt_block(nr, xo, gf, sb, mix, ops.empty_state(), w))
        zero, one = ops.zero_byte(), ops.byte_constant(1)
        j0 = self.nonce + [zero, zero, zero, one]
        self.ct = gctr_target(builder, ops, nr, (xo, gf, sb), mix, w, inc32_target(builder, j0), self.pt)
        if not tag:
            return
        u = (-L) % 16
        a_len = [ops.byte_constant(v) for v in (0).to_bytes(8, "big")]
        c_len = [ops.byte_constant(v) for v in (L * 8).to_bytes(8, "big")]
        ghash_in = self.ct + [zero] * u + a_len + c_len
        sh, br = u8_unit_right_shift_lut(builder), u8_bitref_lut(builder)
        s = ghash_target(builder, xo, sh, br, h, ghash_in)
        t = gctr_target(builder, ops, nr, (xo, gf, sb), mix, w, j0, s)[:TAG_LEN // 8]
        for a, c in zip(self.tag, t):
            builder.connect(a, c)

    def input_targets(self):
        return self.key + self.nonce + self.pt + self.ct + self.tag

    def input_values(self, key, nonce, pt, ct, tag):
        assert len(pt) == self.L and len(ct) == self.L   # circuit_gcm.rs:189-193 copy_from_slice
        tagv = list(tag) if self.with_tag else [0] * (TAG_LEN // 8)
        return list(key) + list(nonce) + list(pt) + list(ct) + tagv

    def set_targets(self, pw, key, nonce, pt, ct, tag):
        for t, v in zip(self.input_targets(), self.input_values(key, nonce, pt, ct, tag)):
            pw.set_target(t, v)
