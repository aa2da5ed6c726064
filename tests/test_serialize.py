"""Wire formats (data_structures.rs:8,31,87,125): ark-serialize CanonicalSerialize restated in
groth16_b200/serialize.py.  BLS12-381 is pinned by the IETF / zcash generator encodings (the format ark-bls12-381
implements); the generic short-Weierstrass format (BN254, BLS12-377) is checked for round trips and flag semantics only."""
import io

import pytest

import pyref as P
from groth16_b200.serialize import ArkCodec
from util import toxic

# standard BLS12-381 generators (IETF pairing-friendly-curves draft / zcash)
G1X = 0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb
G1Y = 0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1
G2X = (0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
       0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e)
G2Y = (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
       0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be)


def test_bls12_381_generator_encodings():
    c = P.CURVES["bls12_381"]
    cx = P.ctx(c)
    g1, g2 = (G1X, G1Y), (G2X, G2Y)
    assert cx.G1.on_curve(g1) and cx.G1.mul(g1, c.r) is None
    assert cx.G2.on_curve(g2) and cx.G2.mul(g2, c.r) is None
    k = ArkCodec("bls12_381")
    assert k.point(g1).hex() == ("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
                                 "6c55e83ff97a1aeffb3af00adb22c6bb")
    assert k.point(g2, g2=True).hex() == ("93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049"
                                          "334cf11213945d57e5ac7d055d042b7e024aa2b2f08f0a91260805272dc51051"
                                          "c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8")
    assert k.point(None).hex() == "c0" + "00" * 47 and k.point(None, g2=True).hex() == "c0" + "00" * 95
    assert k.point(None, compress=False).hex() == "40" + "00" * 95
    # the negated generator only flips the sort flag
    neg = cx.G1.neg(g1)
    assert k.point(neg).hex()[0] == "b" and k.point(neg).hex()[1:] == k.point(g1).hex()[1:]
    assert k.point(g1, compress=False).hex() == "%096x%096x" % (G1X, G1Y)


@pytest.mark.parametrize("curve", ["bls12_381", "bn254", "bls12_377"])
def test_roundtrip_proof_and_keys(curve):
    c = P.CURVES[curve]
    cx = P.ctx(c)
    k = ArkCodec(curve)
    rng = P.Rng(5)
    cs = P.silly_circuit(c, rng.fr(c.r), rng.fr(c.r))
    pk = P.generate_parameters(cs, *toxic(c, 9))
    pf = P.create_proof(pk, cs, rng.fr(c.r), rng.fr(c.r))
    for compress in (True, False):
        data = k.proof(pf.a, pf.b, pf.c, compress)
        nb = k.fq_bytes
        assert len(data) == (4 * nb if compress else 8 * nb)          # a: 1|2, b: 2|4, c: 1|2 field elements
        assert k.read_proof(data, compress) == (pf.a, pf.b, pf.c)
        vk = (pk.vk.alpha_g1, pk.vk.beta_g2, pk.vk.gamma_g2, pk.vk.delta_g2, pk.vk.gamma_abc_g1)
        blob = k.proving_key(vk, pk.beta_g1, pk.delta_g1, pk.a_query, pk.b_g1_query, pk.b_g2_query, pk.h_query, pk.l_query, compress)
        got = k.read_proving_key(blob, compress)
        assert got == (vk, pk.beta_g1, pk.delta_g1, pk.a_query, pk.b_g1_query, pk.b_g2_query, pk.h_query, pk.l_query)
    # identity elements inside queries survive (b_g1_query of this circuit has them)
    assert any(p is None for p in pk.b_g1_query)


def test_generic_sw_flags():
    """ark-ec SWFlags: bit 7 of the last byte <=> y > -y, bit 6 <=> infinity; x little-endian."""
    c = P.CURVES["bn254"]
    k = ArkCodec("bn254")
    g = (1, 2)                                 # BN254 G1 generator
    enc = k.point(g)
    assert len(enc) == 32 and enc[0] == 1 and enc[-1] & 0xC0 == 0      # y = 2 <= -y : "positive"
    neg = (1, c.q - 2)
    assert k.point(neg)[-1] & 0x80 and k.point(neg)[:-1] == enc[:-1]
    assert k.point(None) == bytes(31) + b"\x40"
    assert k.read_point(io.BytesIO(k.point(neg))) == neg
    assert k.fr(5) == (5).to_bytes(32, "little")
