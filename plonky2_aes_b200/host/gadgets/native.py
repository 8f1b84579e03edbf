"""Byte-level AES (FIPS-197) and AES-GCM (SP 800-38D) used as the witness source, mirroring
the structure of the reference's native implementation (/root/reference/aes-gcm/src/
native_aes.rs:27-150, native_gcm.rs:16-250).  Host-side; validated in tests against the FIPS /
NIST CAVP vectors the reference's own tests use and against the `cryptography` package."""


def gf_2_8_mul(a, b):
    """GF(2^8) product modulo x^8+x^4+x^3+x+1 (FIPS-197 §4.2; native_aes.rs:82)."""
    r = 0
    for _ in range(8):
        if b & 1:
            r ^= a
        hi = a & 0x80
        a = (a << 1) & 0xFF
        if hi:
            a ^= 0x1B
        b >>= 1
    return r


def _make_sbox():
    # multiplicative inverse followed by the affine map of FIPS-197 §5.1.1
    inv = [0] * 256
    for x in range(1, 256):
        for y in range(1, 256):
            if gf_2_8_mul(x, y) == 1:
                inv[x] = y
                break
    box = []
    for x in range(256):
        b = inv[x]
        r = 0
        for i in range(8):
            bit = ((b >> i) ^ (b >> ((i + 4) % 8)) ^ (b >> ((i + 5) % 8)) ^ (b >> ((i + 6) % 8)) ^ (b >> ((i + 7) % 8)) ^ (0x63 >> i)) & 1
            r |= bit << i
        box.append(r)
    return box


SBOX = _make_sbox()
RCON = [0x00, 0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1B, 0x36]
TAG_LEN = 128


def rot_word(w):
    return [w[(i + 1) % 4] for i in range(4)]


def shift_rows(s):
    return [[s[i][(i + j) % 4] for j in range(4)] for i in range(4)]


def key_expansion(key, nk, nr):
    w = [list(key[4 * i:4 * i + 4]) for i in range(nk)]
    for i in range(nk, 4 * (nr + 1)):
        temp = w[i - 1]
        if i % nk == 0:
            temp = [SBOX[b] for b in rot_word(temp)]
            temp[0] ^= RCON[i // nk]
        elif nk > 6 and i % nk == 4:
            temp = [SBOX[b] for b in temp]
        w.append([a ^ b for a, b in zip(w[i - nk], temp)])
    return w


def mix_columns(s):
    r = [[0] * 4 for _ in range(4)]
    m = [[2, 3, 1, 1], [1, 2, 3, 1], [1, 1, 2, 3], [3, 1, 1, 2]]
    for c in range(4):
        for i in range(4):
            v = 0
            for k in range(4):
                v ^= gf_2_8_mul(m[i][k], s[k][c])
            r[i][c] = v
    return r


def encrypt_block(block, w, nr):
    """returns the state s[row][col]; flatten with flatten_state."""
    s = [[block[i + 4 * j] for j in range(4)] for i in range(4)]

    def ark(s, rk):
        return [[s[i][j] ^ rk[j][i] for j in range(4)] for i in range(4)]
    s = ark(s, w[0:4])
    for rnd in range(1, nr):
        s = [[SBOX[b] for b in row] for row in s]
        s = shift_rows(s)
        s = mix_columns(s)
        s = ark(s, w[4 * rnd:4 * rnd + 4])
    s = [[SBOX[b] for b in row] for row in s]
    s = shift_rows(s)
    return ark(s, w[4 * nr:4 * nr + 4])


def flatten_state(s):
    return [s[i % 4][i // 4] for i in range(16)]


def inc32(block):
    r = list(block)
    ctr = (int.from_bytes(bytes(block[12:16]), "big") + 1) & 0xFFFFFFFF
    r[12:16] = list(ctr.to_bytes(4, "big"))
    return r


def gctr(w, nr, icb, x):
    y, cb = [], list(icb)
    for i in range(0, len(x), 16):
        if i > 0:
            cb = inc32(cb)
        ks = flatten_state(encrypt_block(cb, w, nr))
        y += [a ^ b for a, b in zip(x[i:i + 16], ks)]
    return y


def gf_2_128_mul(x, y):
    """bitwise GF(2^128) multiplication exactly as the circuit does it (native_gcm.rs:161)."""
    z, v = [0] * 16, list(y)
    for i in range(128):
        if (x[i // 8] >> (7 - i % 8)) & 1:
            z = [a ^ b for a, b in zip(z, v)]
        lsb = v[15] & 1
        carry, nv = 0, []
        for b in v:
            nv.append((carry << 7) | (b >> 1))
            carry = b & 1
        v = nv
        if lsb:
            v[0] ^= 225
    return z


def ghash(h, x):
    y = [0] * 16
    for i in range(0, len(x), 16):
        y = gf_2_128_mul([a ^ b for a, b in zip(y, x[i:i + 16])], h)
    return y


def gcm_encrypt(key, nonce, pt, nk=4, nr=10, aad=b""):
    """AES-GCM with a 96-bit nonce (native_gcm.rs:16-68): returns (ciphertext, tag)."""
    key, nonce, pt, aad = list(key), list(nonce), list(pt), list(aad)
    w = key_expansion(key, nk, nr)
    h = flatten_state(encrypt_block([0] * 16, w, nr))
    j0 = nonce + [0, 0, 0, 1]
    ct = gctr(w, nr, inc32(j0), pt)
    u = (-len(ct)) % 16
    v = (-len(aad)) % 16
    s_in = aad + [0] * v + ct + [0] * u + list((len(aad) * 8).to_bytes(8, "big")) + list((len(ct) * 8).to_bytes(8, "big"))
    s = ghash(h, s_in)
    tag = gctr(w, nr, j0, s)[:TAG_LEN // 8]
    return bytes(ct), bytes(tag)
