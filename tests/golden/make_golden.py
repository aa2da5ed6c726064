#!/usr/bin/env python3
"""Generate tests/golden/golden.json.

The reference (/root/reference) holds no golden vectors and cannot be run here (Rust, un-vendored crates), so these are
NOT reference outputs.  Two kinds of entries:
  * "external": published constants every correct implementation must reproduce (IETF / zcash BLS12-381 generator
    encodings; the SURVEY.md section 2b field constants).
  * "oracle": outputs of oracle/pyref.py (pure big-int restatement, pairing-checked) on fixed seeds -- regression vectors
    that pin the C++ oracle and the CUDA path to the same bits across rounds: domain generators, a 16-point NTT, an
    8-point MSM in G1 and G2, and a complete proof of MySillyCircuit (src/test.rs:14-43) per curve, in the ark-serialize
    compressed wire format of groth16_b200/serialize.py.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import pyref as P  # noqa: E402
from groth16_b200.serialize import ArkCodec  # noqa: E402

SEEDS = dict(toxic=101, witness=202, rs=303, ntt=404, msm=505)


def main():
    out = {"external": {
        "bls12_381_g1_generator_compressed": "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb",
        "bls12_381_g2_generator_compressed": "93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"
                                             "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8",
        "mont_inv64": {"bls12_381": ["0xfffffffeffffffff", "0x89f3fffcfffcfffd"], "bn254": ["0xc2e1f593efffffff", "0x87d20782e4866389"],
                       "bls12_377": ["0x0a117fffffffffff", "0x8508bfffffffffff"]},
        # alt_bn128 / BN254: generator (1, 2) and its double as published with EIP-196 (ecAdd / ecMul test vectors), the G2
        # generator of EIP-197; the G1 generator of ark-bls12-377 (curves/g1.rs G1_GENERATOR_X / _Y)
        "bn254_g1_generator": ["0x1", "0x2"],
        "bn254_two_g1_eip196": ["0x030644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd3",
                                "0x15ed738c0e0a7c92e7845f96b2ae9c0a68a6a449e3538fc7ff3ebf7a5a18a2c4"],
        "bn254_g2_generator_eip197": [["0x1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed",
                                       "0x198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2"],
                                      ["0x12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa",
                                       "0x090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b"]],
        "bls12_377_g1_generator_ark": ["0x008848defe740a67c8fc6225bf87ff5485951e2caa9d41bb188282c8bd37cb5cd5481512ffcd394eeab9b16eb21be9ef",
                                       "0x01914a69c5102eff1f674f5d30afeec4bd7fb348ca3e52d96d182ad44fb82305c2fe3d3634a9591afd82de55559c8ea6"],
    }, "oracle": {}, "seeds": SEEDS}
    for name, c in P.CURVES.items():
        cx = P.ctx(c)
        k = ArkCodec(name)
        e = {}
        e["domain_generator_log"] = {str(L): hex(P.Domain(c, 1 << L).omega) for L in (1, 4, 10, 20)}
        rng = P.Rng(SEEDS["ntt"])
        vals = [rng.fr(c.r) for _ in range(16)]
        dom = P.Domain(c, 16)
        e["ntt16"] = {"in": [hex(v) for v in vals], "fft": [hex(v) for v in dom.fft(vals)],
                      "coset_ifft": [hex(v) for v in dom.ifft(vals, offset=c.fr_gen)]}
        rng = P.Rng(SEEDS["msm"])
        sc = [rng.fr(c.r) for _ in range(8)]
        sc[0], sc[1] = 0, c.r - 1
        b1 = [cx.G1.mul(cx.g1_gen(), rng.fr(c.r)) for _ in range(8)]
        b2 = [cx.G2.mul(cx.g2_gen(), rng.fr(c.r)) for _ in range(8)]
        b1[2] = None
        e["msm8"] = {"scalars": [hex(v) for v in sc], "g1_bases": [k.point(p).hex() for p in b1],
                     "g2_bases": [k.point(p, g2=True).hex() for p in b2],
                     "g1_result": k.point(cx.G1.msm_naive(b1, sc)).hex(), "g2_result": k.point(cx.G2.msm_naive(b2, sc), g2=True).hex()}
        rng = P.Rng(SEEDS["witness"])
        a, b = rng.fr(c.r), rng.fr(c.r)
        cs = P.silly_circuit(c, a, b)
        trng = P.Rng(SEEDS["toxic"])
        tox = [trng.fr(c.r) for _ in range(5)]
        pk = P.generate_parameters(cs, *tox)
        rrng = P.Rng(SEEDS["rs"])
        r_, s_ = rrng.fr(c.r), rrng.fr(c.r)
        pf = P.create_proof(pk, cs, r_, s_)
        assert P.verify_proof(pk.vk, c, pf, [a * b % c.r])
        e["silly_proof"] = {"a": hex(a), "b": hex(b), "toxic": [hex(t) for t in tox], "r": hex(r_), "s": hex(s_),
                            "h": [hex(v) for v in P.witness_map(cs)], "proof_compressed": k.proof(pf.a, pf.b, pf.c).hex(),
                            "vk_alpha_g1": k.point(pk.vk.alpha_g1).hex(), "pk_h_query": [k.point(p).hex() for p in pk.h_query]}
        out["oracle"][name] = e
    with open(os.path.join(ROOT, "tests", "golden", "golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote tests/golden/golden.json")


if __name__ == "__main__":
    main()
