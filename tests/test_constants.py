"""Self-checks of the curve constants and of the generated CUDA table."""
import os
import re
import subprocess
import sys

import pytest

import pyref as P
from groth16_b200.params import CURVES, GENERATORS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _is_probable_prime(n):
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


@pytest.mark.parametrize("name", list(P.CURVES))
def test_field_constants(name):
    c = P.CURVES[name]
    assert _is_probable_prime(c.r) and _is_probable_prime(c.q)
    assert (c.r - 1) % (1 << c.two_adicity) == 0 and ((c.r - 1) >> c.two_adicity) % 2 == 1
    root = pow(c.fr_gen, (c.r - 1) >> c.two_adicity, c.r)
    assert pow(root, 1 << c.two_adicity, c.r) == 1 and pow(root, 1 << (c.two_adicity - 1), c.r) == c.r - 1
    # SURVEY section 2b Montgomery inverses
    want = {"bls12_381": (0xfffffffeffffffff, 0x89f3fffcfffcfffd), "bn254": (0xc2e1f593efffffff, 0x87d20782e4866389),
            "bls12_377": (0x0a117fffffffffff, 0x8508bfffffffffff)}[name]
    assert (P.mont_inv64(c.r), P.mont_inv64(c.q)) == want
    # >= 1 spare top bit in every modulus (the carry-free CIOS shortcut and the 2p < 2^(64N) bound rely on it)
    for p in (c.r, c.q):
        assert p.bit_length() < 64 * ((p.bit_length() + 63) // 64)
    pp = CURVES[name]
    assert (pp.r, pp.q, pp.fr_generator, pp.two_adicity, pp.cid) == (c.r, c.q, c.fr_gen, c.two_adicity, c.cid)


@pytest.mark.parametrize("name", list(P.CURVES))
def test_generators(name):
    c = P.CURVES[name]
    cx = P.ctx(c)
    g1, g2 = GENERATORS[name]["g1"], GENERATORS[name]["g2"]
    assert g1 == cx.g1_gen() and g2 == cx.g2_gen()
    assert cx.G1.on_curve(g1) and cx.G2.on_curve(g2)
    assert cx.G1.mul(g1, c.r) is None and cx.G2.mul(g2, c.r) is None
    if name == "bls12_377":
        assert cx.b2 == (0, 155198655607781456406391640216936120121836107652948796323930557600032281009004493664981332883744016074664192874906)
    if name == "bn254":
        assert cx.b2[0] == 19485874751759354771024239261021720505790618469301721065564631296452457478373


def test_generated_header_is_current_and_correct():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_constants.py")], capture_output=True, text=True, check=True).stdout
    with open(os.path.join(ROOT, "groth16_b200", "csrc", "g16_constants.h")) as f:
        assert f.read().strip() == out.strip()
    # spot-check one table against pyref: BLS12-381 Fr modulus limbs and R mod r
    m = re.search(r"struct BLS381_FrP \{.*?mod\(int i\) \{ constexpr uint32_t t\[8\] = \{([^}]*)\}", out, re.S)
    limbs = [int(x.strip().rstrip("u"), 16) for x in m.group(1).split(",")]
    assert sum(l << (32 * i) for i, l in enumerate(limbs)) == P.BLS12_381.r
