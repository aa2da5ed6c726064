"""GPU probe of one rank of a `world`-way sharded proof on a single GPU: per-phase timings of g16_prove_partial."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from groth16_b200 import Groth16, _lib
from groth16_b200.params import GENERATORS
from groth16_b200.workload import synthetic_r1cs

curve, log_n, world = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
m, z, pub = synthetic_r1cs(curve, log_n, seed=1)
g = Groth16(curve, 0)
G = GENERATORS[g.curve.name]
pk = g.generate_parameters_with_qap(m, 11, 22, 33, 44, 55, G["g1"], G["g2"], export=True)
g.load_proving_key(pk, 0, world)
r = g.codec.fr.enc1(123456789)
out = np.zeros(g.partial_limbs(), dtype=np.uint64)
for flags in (0, _lib.SERIAL_MSMS):
    for i in range(3):
        t = time.time()
        g.prove_partial_raw(r, z.ctypes.data, flags, out)
        wall = (time.time() - t) * 1e3
        tm = g.timings()
        print(f"world={world} flags={flags} wall={wall:.2f}ms total={tm['total_ms']:.2f} h2d={tm['h2d_ms']:.2f} wm={tm['witness_map_ms']:.2f} "
              f"msm={ {k: round(v,2) for k,v in tm['msm_ms'].items()} } accum={ {k: round(v,2) for k,v in tm['msm_accum_ms'].items()} } "
              f"host={tm['host_finish_ms']:.2f} launches={tm['launches']}", flush=True)
