import faulthandler, sys, os, ctypes as C
faulthandler.dump_traceback_later(60, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from plonky2_aes_b200.host.polynomial_batch import Context
from tests import oracle_lib
orc = oracle_lib.load()
ctx = Context(0)
P = 0xFFFFFFFF00000001
rng = np.random.default_rng(3)
for trial, st in enumerate([rng.integers(0, P, size=12, dtype=np.uint64), np.zeros(12, dtype=np.uint64), np.arange(12, dtype=np.uint64)]):
    for bits in (1, 4, 8):
        nonce = C.c_uint64()
        rc = ctx.lib.p2g_pow_grind(ctx.handle, st.ctypes.data, 5, bits, C.byref(nonce))
        exp = None
        for cand in range(100000):
            s = st.copy(); s[5] = cand
            if int(orc.poseidon(s)[7]) >> (64 - bits) == 0:
                exp = cand; break
        print(trial, bits, "rc", rc, "gpu", nonce.value, "oracle", exp, flush=True)
# hash of uniform rows through the generic hash kernel (all threads same input)
rows = np.tile(np.arange(8, dtype=np.uint64), (64, 1))
print(ctx.hash_no_pad_many(rows)[0], orc.hash_no_pad(rows[0]))
