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


def check_prepare_inputs(g, curve_name: str):
    """Groth16::prepare_inputs (verifier.rs:25-39) through `g` against the group law of the big-int oracle; both input
    forms (ints, Montgomery limbs); MalformedVerifyingKey on a length mismatch (verifier.rs:30)."""
    import pytest
    from groth16_b200 import MalformedKey
    c = P.CURVES[curve_name]
    cx = P.ctx(c)
    cd = CurveCodec(get_curve(curve_name))
    rng = P.Rng(77)
    pts = [cx.G1.mul(cx.g1_gen(), rng.fr(c.r)) for _ in range(5)]
    xs = [0, 1, c.r - 1, rng.fr(c.r)]
    vk = VerifyingKey(None, None, None, None, cd.enc_g1(pts))
    want = pts[0]
    for x, b in zip(xs, pts[1:]):
        want = cx.G1.add(want, cx.G1.mul(b, x))
    assert cd.dec_proj_g1(g.prepare_inputs(vk, xs)) == want
    assert cd.dec_proj_g1(g.prepare_inputs(vk, cd.fr.enc(xs))) == want
    with pytest.raises(MalformedKey):
        g.prepare_inputs(vk, xs[:-1])
    with pytest.raises(MalformedKey):
        g.prepare_inputs(VerifyingKey(None, None, None, None, None), xs)


def oracle_setup(curve_name: str, m, toxic_ints, g1=None, g2=None, threads=4):
    """CPU trusted setup by the C++ oracle (orc.generate_parameters) -> (ProvingKey in ABI form, exponent dict)."""
    import orc
    from groth16_b200.params import GENERATORS
    cp = get_curve(curve_name)
    cd = CurveCodec(cp)
    G = GENERATORS[cp.name]
    g1a = cd.enc_g1([g1 or G["g1"]])[0]
    g2a = cd.enc_g2([g2 or G["g2"]])[0]
    k = orc.generate_parameters(cp.cid, cd.nq, m, cd.fr.enc(list(toxic_ints)), g1a, g2a, threads)
    vk = VerifyingKey(k["alpha_g1"], k["beta_g2"], k["gamma_g2"], k["delta_g2"], k["gamma_abc_g1"])
    pk = ProvingKey(vk, k["beta_g1"], k["delta_g1"], k["a_query"], k["b_g1_query"], k["b_g2_query"], k["h_query"], k["l_query"])
    return pk, k["exponents"]
