"""Shared helpers for the parity tests: bridge between oracle/pyref.py objects (Python ints) and the C-ABI arrays."""
import numpy as np

import pyref as P
from groth16_b200 import ConstraintMatrices, CurveCodec, ProvingKey, VerifyingKey, get_curve

ALL_CURVES = ["bls12_381", "bn254", "bls12_377"]


def matrices_from_r1cs(cs: "P.R1CS") -> ConstraintMatrices:
    return ConstraintMatrices.from_rows(cs.curve.name, cs.num_instance, cs.num_witness, cs.a, cs.b, cs.c)


def pk_to_abi(pk: "P.ProvingKey") -> ProvingKey:
    cd = CurveCodec(get_curve(pk.curve.name))
    vk = VerifyingKey(cd.enc_g1([pk.vk.alpha_g1])[0], cd.enc_g2([pk.vk.beta_g2])[0], cd.enc_g2([pk.vk.gamma_g2])[0],
                      cd.enc_g2([pk.vk.delta_g2])[0], cd.enc_g1(pk.vk.gamma_abc_g1))
    return ProvingKey(vk, cd.enc_g1([pk.beta_g1])[0], cd.enc_g1([pk.delta_g1])[0], cd.enc_g1(pk.a_query),
                      cd.enc_g1(pk.b_g1_query), cd.enc_g2(pk.b_g2_query), cd.enc_g1(pk.h_query), cd.enc_g1(pk.l_query))


def pk_from_abi(curve_name: str, pk: ProvingKey, toxic=None) -> "P.ProvingKey":
    cd = CurveCodec(get_curve(curve_name))
    vk = P.VerifyingKey(cd.dec_g1(pk.vk.alpha_g1)[0], cd.dec_g2(pk.vk.beta_g2)[0], cd.dec_g2(pk.vk.gamma_g2)[0],
                        cd.dec_g2(pk.vk.delta_g2)[0], cd.dec_g1(pk.vk.gamma_abc_g1))
    return P.ProvingKey(P.CURVES[curve_name], vk, cd.dec_g1(pk.beta_g1)[0], cd.dec_g1(pk.delta_g1)[0],
                        cd.dec_g1(pk.a_query), cd.dec_g1(pk.b_g1_query), cd.dec_g2(pk.b_g2_query),
                        cd.dec_g1(pk.h_query), cd.dec_g1(pk.l_query), toxic=toxic)


def proof_from_abi(curve_name: str, pf) -> "P.Proof":
    cd = CurveCodec(get_curve(curve_name))
    return P.Proof(cd.dec_g1(pf.a)[0], cd.dec_g2(pf.b)[0], cd.dec_g1(pf.c)[0])


def toxic(curve, seed):
    rng = P.Rng(seed)
    return [rng.fr(curve.r) for _ in range(5)]
