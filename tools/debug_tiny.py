import faulthandler, sys, os
faulthandler.dump_traceback_later(40, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonky2_aes_b200.host.polynomial_batch import Context
from tests import circuits
ctx = Context(0)
print("ctx ok", flush=True)
data, wires = circuits.tiny_arith()
print("built", flush=True)
data.load(ctx)
print("loaded", flush=True)
p = data.prove_wires(wires)
print("proved", len(p), flush=True)
data2, wires2, _ = circuits.aes_gcm(13, True)
data2.load(ctx)
print("loaded2", flush=True)
p = data2.prove_wires(wires2)
print("proved2", len(p), flush=True)
