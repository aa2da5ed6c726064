"""GPU probe: setup + prove a synthetic circuit of 2^log_n and print the per-phase timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from groth16_b200 import Groth16
from groth16_b200.params import GENERATORS
from groth16_b200.workload import synthetic_r1cs, dummy_r1cs
from groth16_b200 import _lib

curve = sys.argv[1] if len(sys.argv) > 1 else "bls12_381"
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
kind = sys.argv[3] if len(sys.argv) > 3 else "synthetic"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
t = time.time()
if kind == "dummy":
    m, z, pub = dummy_r1cs(curve, (1 << log_n) - 100, (1 << log_n) - 100)
else:
    m, z, pub = synthetic_r1cs(curve, log_n, seed=1)
print(f"workload {time.time()-t:.1f}s nc={m.num_constraints} nw={m.num_witness_variables}", flush=True)
g = Groth16(curve, 0)
t = time.time()
G = GENERATORS[g.curve.name]
g.generate_parameters_with_qap(m, 11, 22, 33, 44, 55, G["g1"], G["g2"], export=False)
print(f"gpu setup {time.time()-t:.1f}s", flush=True)
r = g.codec.fr.enc1(123456789); s = g.codec.fr.enc1(987654321)
out = np.zeros(8 * g.nq, dtype=np.uint64)
for flags in (0, _lib.SERIAL_MSMS):
    for i in range(reps):
        t = time.time()
        g.prove_raw(r, s, z.ctypes.data, flags, out)
        wall = (time.time() - t) * 1e3
        tm = g.timings()
        print(f"flags={flags} wall={wall:.2f}ms total={tm['total_ms']:.2f} h2d={tm['h2d_ms']:.2f} wm={tm['witness_map_ms']:.2f} "
              f"msm={ {k: round(v,2) for k,v in tm['msm_ms'].items()} } accum={ {k: round(v,2) for k,v in tm['msm_accum_ms'].items()} } "
              f"host={tm['host_finish_ms']:.2f} launches={tm['launches']}", flush=True)
print("proof a[0:2]", out[:2])
