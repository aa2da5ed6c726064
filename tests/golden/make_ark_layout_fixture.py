#!/usr/bin/env python3
"""Writes fixtures in the directory layout of oracle/ark_fixture (the Rust program that runs REAL ark-groth16), but produced
by this repository's own oracle -- so that tests/test_ark_fixture.py (loader, codecs, prove-and-compare) is exercised even
though no Rust toolchain exists in the build image.  Every meta.json says so: "producer": "oracle ... NOT arkworks".
A directory written by the Rust program has "producer": "ark-groth16 0.5.0 ..." and is what actually pins parity.

    python tests/golden/make_ark_layout_fixture.py        (rewrites tests/golden/ark/oracle_*)
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import orc  # noqa: E402
import pyref as P  # noqa: E402
from groth16_b200 import CurveCodec, get_curve  # noqa: E402
from groth16_b200.serialize import ArkCodec  # noqa: E402
from util import matrices_from_r1cs, oracle_setup, pk_from_abi, toxic  # noqa: E402


def write(curve, name, cs, seed):
    c = P.CURVES[curve]
    cd = CurveCodec(get_curve(curve))
    k = ArkCodec(curve)
    rng = P.Rng(seed)
    m = matrices_from_r1cs(cs)
    cx = P.ctx(c)
    pk_abi, _ = oracle_setup(curve, m, toxic(c, seed), cx.g1_gen(), cx.g2_gen(), threads=4)
    pk = pk_from_abi(curve, pk_abi)
    r_, s_ = rng.fr(c.r), rng.fr(c.r)
    z = cd.fr.enc(cs.assignment)
    proof, _ = orc.prove(c.cid, cd.nq, pk_abi, m, z, cd.fr.enc1(r_), cd.fr.enc1(s_), threads=4)
    nq = cd.nq
    a, b, cc = cd.dec_g1(proof[:2 * nq])[0], cd.dec_g2(proof[2 * nq:6 * nq])[0], cd.dec_g1(proof[6 * nq:])[0]
    h = cd.fr.dec(orc.witness_map(c.cid, m, z, threads=2))
    d = os.path.join(HERE, "ark", f"oracle_{curve}_{name}")
    os.makedirs(d, exist_ok=True)
    vk = (pk.vk.alpha_g1, pk.vk.beta_g2, pk.vk.gamma_g2, pk.vk.delta_g2, pk.vk.gamma_abc_g1)
    files = {
        "pk.bin": k.proving_key(vk, pk.beta_g1, pk.delta_g1, pk.a_query, pk.b_g1_query, pk.b_g2_query, pk.h_query, pk.l_query, compress=False),
        "vk.bin": k.verifying_key(*vk, compress=True),
        "proof.bin": k.proof(a, b, cc, compress=True),
        "proof_uncompressed.bin": k.proof(a, b, cc, compress=False),
        "matrices.bin": k.matrices(cs.num_instance, cs.num_witness, cs.a, cs.b, cs.c),
        "witness.bin": k.fr_vec(cs.assignment),
        "public.bin": k.fr_vec(cs.assignment[1:cs.num_instance]),
        "rs.bin": k.fr_vec([r_, s_]),
        "h.bin": k.fr_vec(h),
    }
    for fn, data in files.items():
        with open(os.path.join(d, fn), "wb") as f:
            f.write(data)
    meta = {"producer": "oracle (oracle/oracle.cpp via tests/golden/make_ark_layout_fixture.py) -- NOT arkworks", "curve": curve,
            "circuit": name, "seed": seed, "num_instance_variables": cs.num_instance, "num_witness_variables": cs.num_witness,
            "num_constraints": cs.num_constraints, "pk": "uncompressed", "proof": "compressed"}
    with open(os.path.join(d, "meta.json"), "w") as f:
        json.dump(meta, f)
        f.write("\n")
    print("wrote", d, sum(len(v) for v in files.values()), "bytes")


def main():
    for curve in ("bls12_381", "bn254", "bls12_377"):
        c = P.CURVES[curve]
        rng = P.Rng(7)
        write(curve, "silly", P.silly_circuit(c, rng.fr(c.r), rng.fr(c.r)), 21)
    c = P.CURVES["bn254"]
    write("bn254", "synthetic_2p6", P.synthetic_circuit(c, 62, seed=4, num_inputs=1), 22)


if __name__ == "__main__":
    main()
