"""CPU tests: the fast C++ oracle (oracle/oracle.cpp) pinned against the big-int reference (oracle/pyref.py), and
pyref's own known-answer checks (the reference /root/reference holds no golden vectors)."""
import numpy as np
import pytest

import orc
import pyref as P
from groth16_b200 import CurveCodec, get_curve
from groth16_b200.workload import dummy_r1cs, synthetic_r1cs
from util import ALL_CURVES, matrices_from_r1cs, pk_to_abi, proof_from_abi, toxic


@pytest.mark.parametrize("curve", ALL_CURVES)
def test_pyref_ntt_vs_naive_dft(curve):
    c = P.CURVES[curve]
    rng = P.Rng(1)
    for log_n in (0, 1, 3, 5):
        d = P.Domain(c, 1 << log_n)
        co = [rng.fr(c.r) for _ in range(d.n)]
        assert d.fft(co) == d.dft_naive(co)
        assert d.fft(co, offset=c.fr_gen) == d.dft_naive(co, offset=c.fr_gen)
        assert d.ifft(d.fft(co)) == co
        assert d.ifft(d.fft(co, offset=c.fr_gen), offset=c.fr_gen) == co
    assert pow(d.omega, d.n, c.r) == 1 and (d.n == 1 or pow(d.omega, d.n // 2, c.r) == c.r - 1)


@pytest.mark.parametrize("curve", ALL_CURVES)
def test_pyref_h_times_z_identity(curve):
    """h(X) Z(X) == A(X) B(X) - C(X) at a random point (SURVEY.md section 8c check 2)."""
    c = P.CURVES[curve]
    rng = P.Rng(2)
    cs = P.synthetic_circuit(c, 13, seed=3, num_inputs=2)
    assert cs.is_satisfied()
    dom, a, b, cc = P.abc_evals(cs)
    h = P.witness_map_from_evals(dom, a, b, cc)
    x = rng.fr(c.r)
    ev = lambda coeffs: sum(cf * pow(x, i, c.r) for i, cf in enumerate(coeffs)) % c.r
    A, B, Cc = ev(dom.ifft(a)), ev(dom.ifft(b)), ev(dom.ifft(cc))
    assert ev(h) * dom.vanishing(x) % c.r == (A * B - Cc) % c.r
    assert h[-1] == 0  # top coefficient vanishes


@pytest.mark.parametrize("curve", ALL_CURVES)
def test_pyref_groth16_roundtrip(curve):
    """src/test.rs:45-73: prove verifies, wrong input rejected; prover == closed form in the exponent; pairing is
    bilinear and non-degenerate."""
    c = P.CURVES[curve]
    cx = P.ctx(c)
    rng = P.Rng(4)
    a, b = rng.fr(c.r), rng.fr(c.r)
    cs = P.silly_circuit(c, a, b)
    pk = P.generate_parameters(cs, *toxic(c, 5))
    r_, s_ = rng.fr(c.r), rng.fr(c.r)
    pf = P.create_proof(pk, cs, r_, s_)
    pe = P.proof_in_the_exponent(pk, cs, r_, s_)
    assert (pf.a, pf.b, pf.c) == (pe.a, pe.b, pe.c)
    assert P.verify_proof(pk.vk, c, pf, [a * b % c.r])
    assert not P.verify_proof(pk.vk, c, pf, [a])
    with pytest.raises(ValueError):
        P.verify_proof(pk.vk, c, pf, [])  # MalformedVerifyingKey, verifier.rs:29-31
    # bilinearity: e(2P, 3Q) e(-6P, Q) == 1 and e(P, Q) != 1
    g1, g2 = cx.g1_gen(), cx.g2_gen()
    assert cx.pairing_product_is_one([(cx.G1.mul(g1, 2), cx.G2.mul(g2, 3)), (cx.G1.mul(g1, c.r - 6), g2)])
    assert not cx.pairing_product_is_one([(g1, g2)])


@pytest.mark.parametrize("curve", ALL_CURVES)
@pytest.mark.parametrize("log_n", [0, 1, 4, 9])
def test_oracle_ntt(curve, log_n):
    c = P.CURVES[curve]
    cd = CurveCodec(get_curve(curve))
    rng = P.Rng(10 + log_n)
    n = 1 << log_n
    vals = [rng.fr(c.r) for _ in range(n)]
    dom = P.Domain(c, n)
    enc = cd.fr.enc(vals)
    for thr in (1, 3):
        assert cd.fr.dec(orc.ntt(c.cid, log_n, enc, threads=thr)) == dom.fft(vals)
        assert cd.fr.dec(orc.ntt(c.cid, log_n, enc, inverse=True, threads=thr)) == dom.ifft(vals)
        assert cd.fr.dec(orc.ntt(c.cid, log_n, enc, coset=True, threads=thr)) == dom.fft(vals, offset=c.fr_gen)
        assert cd.fr.dec(orc.ntt(c.cid, log_n, enc, inverse=True, coset=True, threads=thr)) == dom.ifft(vals, offset=c.fr_gen)


@pytest.mark.parametrize("curve", ALL_CURVES)
def test_oracle_msm(curve):
    c = P.CURVES[curve]
    cx = P.ctx(c)
    cd = CurveCodec(get_curve(curve))
    rng = P.Rng(20)
    n = 70
    b1 = [cx.G1.mul(cx.g1_gen(), rng.fr(c.r)) for _ in range(n)]
    b2 = [cx.G2.mul(cx.g2_gen(), rng.fr(c.r)) for _ in range(24)]
    sc = [rng.fr(c.r) for _ in range(n)]
    sc[:6] = [0, 1, c.r - 1, 2, (1 << 200) % c.r, 5]
    b1[3] = None
    b1[5] = b1[4]; sc[5] = sc[4]
    b1[7] = cx.G1.neg(b1[6]); sc[7] = sc[6]
    for thr in (1, 4):
        assert cd.dec_proj_g1(orc.msm_g1(c.cid, cd.nq, cd.enc_g1(b1), cd.fr.bigint(sc), threads=thr)) == cx.G1.msm_naive(b1, sc)
    assert cd.dec_proj_g2(orc.msm_g2(c.cid, cd.nq, cd.enc_g2(b2), cd.fr.bigint(sc), threads=2)) == cx.G2.msm_naive(b2, sc)
    assert cd.dec_proj_g1(orc.msm_g1(c.cid, cd.nq, cd.enc_g1(b1[:0]), cd.fr.bigint(sc[:0]))) is None


@pytest.mark.parametrize("curve", ALL_CURVES)
def test_oracle_prove_matches_pyref(curve):
    c = P.CURVES[curve]
    cd = CurveCodec(get_curve(curve))
    rng = P.Rng(30)
    cs = P.synthetic_circuit(c, 21, seed=6, num_inputs=2)
    opk = P.generate_parameters(cs, *toxic(c, 7))
    m = matrices_from_r1cs(cs)
    z = cd.fr.enc(cs.assignment)
    assert cd.fr.dec(orc.witness_map(c.cid, m, z, threads=2)) == P.witness_map(cs)
    for r_, s_ in ((rng.fr(c.r), rng.fr(c.r)), (0, rng.fr(c.r))):
        proof, _ = orc.prove(c.cid, cd.nq, pk_to_abi(opk), m, z, cd.fr.enc1(r_), cd.fr.enc1(s_), threads=3)
        nq = cd.nq
        from groth16_b200 import Proof
        pf = proof_from_abi(curve, Proof(proof[:2 * nq], proof[2 * nq:6 * nq], proof[6 * nq:]))
        want = P.create_proof(opk, cs, r_, s_)
        assert (pf.a, pf.b, pf.c) == (want.a, want.b, want.c)


def test_oracle_mimc_bls12_377():
    """tests/mimc.rs (BLS12-377, 644 constraints, domain 2^10): witness map and proof vs the closed form."""
    c = P.CURVES["bls12_377"]
    cd = CurveCodec(get_curve("bls12_377"))
    rng = P.Rng(40)
    constants = [rng.fr(c.r) for _ in range(P.MIMC_ROUNDS)]
    cs = P.mimc_circuit(c, rng.fr(c.r), rng.fr(c.r), constants)
    assert cs.is_satisfied() and cs.num_constraints == 644 and cs.num_witness == 645
    m = matrices_from_r1cs(cs)
    z = cd.fr.enc(cs.assignment)
    assert cd.fr.dec(orc.witness_map(c.cid, m, z, threads=4)) == P.witness_map(cs)


@pytest.mark.parametrize("curve", ALL_CURVES)
def test_workload_generators_are_satisfiable(curve):
    """groth16_b200.workload (bench inputs): satisfiable, exact power-of-two domain, DummyCircuit shape."""
    c = P.CURVES[curve]
    cd = CurveCodec(get_curve(curve))
    m, z, pub = synthetic_r1cs(curve, 6, seed=3)
    assert m.num_constraints + m.num_instance_variables == 64
    zi = cd.fr.dec(z)

    def rows(t):
        rp, col, val = t
        v = cd.fr.dec(val)
        return [[(v[e], int(col[e])) for e in range(rp[i], rp[i + 1])] for i in range(len(rp) - 1)]

    cs = P.R1CS(c, m.num_instance_variables, m.num_witness_variables, rows(m.a), rows(m.b), rows(m.c), zi)
    assert cs.is_satisfied() and zi[0] == 1 and zi[1:m.num_instance_variables] == pub
    m2, z2, pub2 = dummy_r1cs(curve, 40, 40)
    z2i = cd.fr.dec(z2)
    cs2 = P.R1CS(c, 2, m2.num_witness_variables, rows(m2.a), rows(m2.b), rows(m2.c), z2i)
    ref = P.dummy_circuit(c, z2i[2], z2i[3], 40, 40)
    assert cs2.is_satisfied() and cs2.a == ref.a and cs2.b == ref.b and cs2.c == ref.c and z2i == ref.assignment


@pytest.mark.parametrize("curve", ALL_CURVES)
def test_oracle_batch_mul(curve):
    """fixed-base batch multiplication (generator.rs:129-183) vs scalar multiplication in pyref."""
    c = P.CURVES[curve]
    cx = P.ctx(c)
    cd = CurveCodec(get_curve(curve))
    rng = P.Rng(50)
    sc = [0, 1, c.r - 1] + [rng.fr(c.r) for _ in range(5)]
    g1 = cd.enc_g1([cx.g1_gen()])[0]
    g2 = cd.enc_g2([cx.g2_gen()])[0]
    assert cd.dec_g1(orc.batch_mul_g1(c.cid, cd.nq, g1, cd.fr.enc(sc), threads=2)) == [cx.G1.mul(cx.g1_gen(), k) for k in sc]
    assert cd.dec_g2(orc.batch_mul_g2(c.cid, cd.nq, g2, cd.fr.enc(sc), threads=2)) == [cx.G2.mul(cx.g2_gen(), k) for k in sc]


@pytest.mark.parametrize("curve", ALL_CURVES)
def test_oracle_setup_matches_pyref(curve):
    """orc.generate_parameters (generator.rs:47-208 restated in C++) against the big-int setup of pyref: every exponent and
    every key element, then a proof under that key verifies with the pairing."""
    from util import oracle_setup
    c = P.CURVES[curve]
    cd = CurveCodec(get_curve(curve))
    cs = P.synthetic_circuit(c, 13, seed=9, num_inputs=2)
    tw = toxic(c, 11)
    cx = P.ctx(c)
    opk = P.generate_parameters(cs, *tw)
    want = P.generate_parameters(cs, *tw, scalars_only=True)
    m = matrices_from_r1cs(cs)
    pk, ex = oracle_setup(curve, m, tw, cx.g1_gen(), cx.g2_gen(), threads=3)
    assert cd.fr.dec(ex["a"]) == want["a"] and cd.fr.dec(ex["b"]) == want["b"]
    assert cd.fr.dec(ex["l"]) == want["l"] and cd.fr.dec(ex["h"]) == want["h"] and cd.fr.dec(ex["gamma_abc"]) == want["gamma_abc"]
    ref = pk_to_abi(opk)
    for name in ("a_query", "b_g1_query", "b_g2_query", "h_query", "l_query", "beta_g1", "delta_g1"):
        assert np.array_equal(np.asarray(getattr(pk, name)).reshape(-1), np.asarray(getattr(ref, name)).reshape(-1)), name
    for name in ("alpha_g1", "beta_g2", "gamma_g2", "delta_g2", "gamma_abc_g1"):
        assert np.array_equal(np.asarray(getattr(pk.vk, name)).reshape(-1), np.asarray(getattr(ref.vk, name)).reshape(-1)), name


@pytest.mark.parametrize("curve", ALL_CURVES)
def test_oracle_window_parallel_msm_agrees(curve):
    """ark-ec's window-parallel MSM schedule and the chunk-parallel default give the same group element."""
    c = P.CURVES[curve]
    cx = P.ctx(c)
    cd = CurveCodec(get_curve(curve))
    rng = P.Rng(61)
    n = 300
    sc = [rng.fr(c.r) for _ in range(n)]
    g1 = cd.enc_g1([cx.g1_gen()])[0]
    bases = orc.batch_mul_g1(c.cid, cd.nq, g1, cd.fr.enc([rng.fr(c.r) for _ in range(n)]), threads=2)
    a = orc.msm_g1(c.cid, cd.nq, bases, cd.fr.bigint(sc), threads=3)
    old = orc.set_msm_mode(1)
    try:
        b = orc.msm_g1(c.cid, cd.nq, bases, cd.fr.bigint(sc), threads=3)
    finally:
        orc.set_msm_mode(old)
    assert np.array_equal(a, b)
