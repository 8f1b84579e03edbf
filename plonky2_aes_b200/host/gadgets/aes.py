"""AES round functions as circuit gadgets over byte targets via lookup tables — the Python
mirror of trait `CircuitBuilderAESState` (/root/reference/aes-gcm/src/circuit_aes.rs:41-274)
and the LUT constructors (:299-358).  Same names, same argument order, same lookup/arithmetic
decomposition: a byte XOR or GF(2^8) product is `256*x + y` followed by one table lookup."""
from .native import SBOX, RCON, gf_2_8_mul as native_gf_2_8_mul, rot_word, shift_rows


def sbox_lut(builder):
    return builder.add_lookup_table_from_pairs([(i, o) for i, o in enumerate(SBOX)])


def byte_xor_lut(builder):
    return builder.add_lookup_table_from_pairs([((x << 8) + y, x ^ y) for x in range(256) for y in range(256)])


def gf_2_8_mul_lut(builder):
    return builder.add_lookup_table_from_pairs(
        [((x << 8) + y, native_gf_2_8_mul(x, y)) for x in range(256) for y in range(256)])


def state_mix_matrix(builder):
    one, two, three = builder.constant(1), builder.constant(2), builder.constant(3)
    return [[two, three, one, one], [one, two, three, one], [one, one, two, three], [three, one, one, two]]


def byte_xor(builder, xor_lut_idx, x, y):
    idx = builder.mul_const_add(1 << 8, x, y)
    return builder.add_lookup_from_index(idx, xor_lut_idx)


def flatten(state):
    return [state[i % 4][i // 4] for i in range(16)]


def from_flat(b):
    return [[b[j * 4 + i] for j in range(4)] for i in range(4)]


class AESStateOps:
    """Methods the reference adds to CircuitBuilder through the CircuitBuilderAESState trait."""

    def __init__(self, builder):
        self.b = builder

    def add_virtual_byte_target_unsafe(self):
        return self.b.add_virtual_target()

    def assert_byte(self, x, u8_table_idx):
        self.b.add_lookup_from_index(x, u8_table_idx)

    def add_virtual_byte_target(self, u8_table_idx):
        t = self.add_virtual_byte_target_unsafe()
        self.assert_byte(t, u8_table_idx)
        return t

    def byte_constant(self, c):
        return self.b.constant(c)

    def zero_byte(self):
        return self.b.zero()

    def empty_state(self):
        z = self.zero_byte()
        return [[z] * 4 for _ in range(4)]

    def add_virtual_state_target(self, u8_table_idx):
        return [[self.add_virtual_byte_target(u8_table_idx) for _ in range(4)] for _ in range(4)]

    def gf_2_8_add(self, xor_lut_idx, x, y):
        return byte_xor(self.b, xor_lut_idx, x, y)

    def gf_2_8_mul(self, gf_2_8_mul_lut_idx, x, y):
        idx = self.b.mul_const_add(1 << 8, x, y)
        return self.b.add_lookup_from_index(idx, gf_2_8_mul_lut_idx)

    def state_sub_word(self, sbox_lut_idx, word):
        return [self.b.add_lookup_from_index(t, sbox_lut_idx) for t in word]

    def state_sub_bytes(self, sbox_lut_idx, s):
        return [self.state_sub_word(sbox_lut_idx, s[i]) for i in range(4)]

    def bytearray_ip(self, xor_lut_idx, gf_lut_idx, x, y):
        acc = self.b.zero()
        for a, c in zip(x, y):
            prod = self.gf_2_8_mul(gf_lut_idx, a, c)
            acc = self.gf_2_8_add(xor_lut_idx, acc, prod)
        return acc

    def byte_matrix_apply(self, xor_lut_idx, gf_lut_idx, a, x):
        return [self.bytearray_ip(xor_lut_idx, gf_lut_idx, a[i], x) for i in range(len(a))]

    def state_mix_columns(self, xor_lut_idx, gf_lut_idx, mix_matrix, s):
        cols = [[s[j][i] for j in range(4)] for i in range(4)]
        out_cols = [self.byte_matrix_apply(xor_lut_idx, gf_lut_idx, mix_matrix, cols[i]) for i in range(4)]
        return [[out_cols[j][i] for j in range(4)] for i in range(4)]

    def state_add_round_key(self, xor_lut_idx, round_key, s):
        return [[self.gf_2_8_add(xor_lut_idx, s[i][j], round_key[j][i]) for j in range(4)] for i in range(4)]

    def key_expansion(self, nk, nr, xor_lut_idx, sbox_lut_idx, key):
        rcon = [self.byte_constant(RCON[i]) for i in range(11)]
        st = [[key[4 * i + j] for j in range(4)] for i in range(nk)]
        for i in range(nk, 4 * (nr + 1)):
            if i % nk == 0:
                term = self.state_sub_word(sbox_lut_idx, rot_word(st[i - 1]))
                offset = [self.gf_2_8_add(xor_lut_idx, term[0], rcon[i // nk])] + term[1:]
            elif nk > 6 and i % nk == 4:
                offset = self.state_sub_word(sbox_lut_idx, st[i - 1])
            else:
                offset = st[i - 1]
            st.append([self.gf_2_8_add(xor_lut_idx, st[i - nk][j], offset[j]) for j in range(4)])
        return st

    def encrypt_block(self, nr, xor_lut_idx, gf_lut_idx, sbox_lut_idx, mix_matrix, s, w):
        s = self.state_add_round_key(xor_lut_idx, w[0:4], s)
        for i in range(1, nr):
            s = self.state_sub_bytes(sbox_lut_idx, s)
            s = shift_rows(s)
            s = self.state_mix_columns(xor_lut_idx, gf_lut_idx, mix_matrix, s)
            s = self.state_add_round_key(xor_lut_idx, w[4 * i:4 * (i + 1)], s)
        s = self.state_sub_bytes(sbox_lut_idx, s)
        s = shift_rows(s)
        return self.state_add_round_key(xor_lut_idx, w[4 * nr:4 * (nr + 1)], s)
