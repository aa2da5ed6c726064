"""Profiling driver (run under ncu with --profile-from-start off): one resident 2^log_n circuit, `warm` untimed proofs,
then ONE proof with the five MSMs serialised inside cudaProfilerStart/Stop, so that the capture holds exactly the kernels
of one proof.  Usage: profile_run.py <curve> <log_n> [opt=value ...]   (options of g16_set_option)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from groth16_b200 import Groth16, _lib
from groth16_b200.params import GENERATORS
from groth16_b200.workload import synthetic_r1cs

curve = sys.argv[1] if len(sys.argv) > 1 else "bls12_381"
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
opts = dict(kv.split("=") for kv in sys.argv[3:])
m, z, pub = synthetic_r1cs(curve, log_n, seed=1)
g = Groth16(curve, 0)
G = GENERATORS[g.curve.name]
g.generate_parameters_with_qap(m, 11, 22, 33, 44, 55, G["g1"], G["g2"], export=False)
for k, v in opts.items():
    g.set_option(k, int(v))
r = g.codec.fr.enc1(123456789); s = g.codec.fr.enc1(987654321)
out = np.zeros(8 * g.nq, dtype=np.uint64)
for i in range(2):
    g.prove_raw(r, s, z.ctypes.data, _lib.SERIAL_MSMS, out)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
g.prove_raw(r, s, z.ctypes.data, _lib.SERIAL_MSMS, out)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(g.timings())
print(g.config())
import json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kernel_rev
tag = os.environ.get("G16_PROFILE_TAG")
if tag:
    json.dump({"kernel_rev": kernel_rev(), "config": g.config(), "curve": curve, "log_n": log_n, "options": opts},
              open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"{tag}_meta.json"), "w"))
