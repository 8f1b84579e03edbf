"""smoke(): one small AES-GCM proof on cuda:0, checked bit for bit against the oracle."""
import numpy as np


def smoke_prove(ctx, orc):
    from tests import circuits, oracle_lib
    data, wires, _ = circuits.aes_gcm(13, True)
    data.load(ctx)
    oc = oracle_lib.OracleCircuit(orc, data)
    g = data.prove_wires(wires)
    o = oc.prove(wires)
    assert np.array_equal(g, o), "GPU proof differs from the oracle proof"
    assert oc.verify(g) == 0, "restated verifier rejected the GPU proof"
    oc.free()
