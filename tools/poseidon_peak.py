import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonky2_aes_b200.host.polynomial_batch import Context
ctx = Context(0)
print("poseidon perms/s", ctx.poseidon_peak(32))
