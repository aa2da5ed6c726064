"""Profiling driver (run under ncu): one resident 2^log_n circuit, `reps` proofs with the five MSMs serialised."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from groth16_b200 import Groth16, _lib
from groth16_b200.params import GENERATORS
from groth16_b200.workload import synthetic_r1cs

curve = sys.argv[1] if len(sys.argv) > 1 else "bls12_381"
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
m, z, pub = synthetic_r1cs(curve, log_n, seed=1)
g = Groth16(curve, 0)
G = GENERATORS[g.curve.name]
g.generate_parameters_with_qap(m, 11, 22, 33, 44, 55, G["g1"], G["g2"], export=False)
r = g.codec.fr.enc1(123456789); s = g.codec.fr.enc1(987654321)
out = np.zeros(8 * g.nq, dtype=np.uint64)
for i in range(reps):
    g.prove_raw(r, s, z.ctypes.data, _lib.SERIAL_MSMS, out)
    print(g.timings())
