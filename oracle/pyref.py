"""oracle/pyref.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Pure-Python big-integer restatement of the Groth16 algebra that ark-groth16 0.5.0 drives
(/root/reference/src/{r1cs_to_qap,generator,prover,verifier}.rs) and of the behaviour of its
un-vendored dependencies (ark-ff / ark-ec / ark-poly 0.5.0, source not on this machine; SURVEY.md
section 2a).  It is the root of trust for the fast C++ oracle (oracle/oracle.cpp) and for the CUDA path.

PARITY PINNING: the reference holds no golden vectors or known-answer tests for this path
(every reference test is a randomized prove->verify round trip) and the reference cannot
be compiled here (no Rust toolchain).  Byte-level parity is therefore "unpinned by reference
fixtures"; it is pinned by mathematics: for fixed (ProvingKey, witness, r, s) a Groth16 proof is a
unique triple of affine points, and this file checks (i) the pairing equation of verifier.rs:44-65,
(ii) the "proof in the exponent" closed form, (iii) NTT vs O(n^2) DFT, (iv) MSM vs double-and-add.

Everything here works on canonical integers; Montgomery form only appears in to_mont/from_mont,
used when packing data for the C ABI.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------------------------
# Curve parameters (self-checked in tests/test_constants.py)
# --------------------------------------------------------------------------------------------


@dataclass(frozen=True)
class Curve:
    name: str
    cid: int  # curve id at the C ABI (include/g16b200.h)
    r: int  # scalar field modulus
    q: int  # base field modulus
    fr_gen: int  # Fr::GENERATOR (multiplicative generator; also the coset offset, r1cs_to_qap.rs:204)
    two_adicity: int
    b: int  # G1: y^2 = x^3 + b
    beta: int  # Fq2 = Fq[u]/(u^2 - beta)
    xi: Tuple[int, int]  # Fq12 = Fq2[w]/(w^6 - xi)
    twist: str  # 'M' (b' = b*xi) or 'D' (b' = b/xi)
    x: int  # curve parameter (signed)
    family: str  # 'bls12' or 'bn'

    @property
    def fr_limbs(self):
        return 4

    @property
    def fq_limbs(self):
        return (self.q.bit_length() + 63) // 64


BLS12_381 = Curve(
    name="bls12_381", cid=0,
    r=0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    q=0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    fr_gen=7, two_adicity=32, b=4, beta=-1, xi=(1, 1), twist="M",
    x=-0xd201000000010000, family="bls12")

BN254 = Curve(
    name="bn254", cid=1,
    r=21888242871839275222246405745257275088548364400416034343698204186575808495617,
    q=21888242871839275222246405745257275088696311157297823662689037894645226208583,
    fr_gen=5, two_adicity=28, b=3, beta=-1, xi=(9, 1), twist="D",
    x=4965661367192848881, family="bn")

BLS12_377 = Curve(
    name="bls12_377", cid=2,
    r=8444461749428370424248824938781546531375899335154063827935233455917409239041,
    q=258664426012969094010652733694893533536393512754914660539884262666720468348340822774968888139573360124440321458177,
    fr_gen=22, two_adicity=47, b=1, beta=-5, xi=(0, 1), twist="D",
    x=0x8508c00000000001, family="bls12")

CURVES = {c.name: c for c in (BLS12_381, BN254, BLS12_377)}
CURVES_BY_ID = {c.cid: c for c in CURVES.values()}


def inv(a: int, p: int) -> int:
    return pow(a, -1, p)


# --------------------------------------------------------------------------------------------
# Montgomery helpers (ark-ff Fp<MontBackend,N>: value stored as a*R mod p, R = 2^(64N); section 2a)
# --------------------------------------------------------------------------------------------


def mont_R(p: int) -> int:
    n = (p.bit_length() + 63) // 64
    return 1 << (64 * n)


def to_mont(a: int, p: int) -> int:
    return (a * mont_R(p)) % p


def from_mont(a: int, p: int) -> int:
    return (a * inv(mont_R(p), p)) % p


def mont_inv64(p: int) -> int:
    """-p^{-1} mod 2^64 (ark-ff MontConfig::INV)."""
    return (-inv(p, 1 << 64)) % (1 << 64)


# --------------------------------------------------------------------------------------------
# Generic field-ops objects so the group law is written once for Fq and Fq2
# --------------------------------------------------------------------------------------------


class FqOps:
    def __init__(self, p):
        self.p = p
        self.zero = 0
        self.one = 1

    def add(self, a, b): return (a + b) % self.p
    def sub(self, a, b): return (a - b) % self.p
    def mul(self, a, b): return (a * b) % self.p
    def neg(self, a): return (-a) % self.p
    def inv(self, a): return pow(a, -1, self.p)
    def is_zero(self, a): return a % self.p == 0
    def eq(self, a, b): return (a - b) % self.p == 0
    def from_int(self, a): return a % self.p

    def sqrt(self, a):
        """Tonelli-Shanks; returns None if a is a non-residue."""
        p = self.p
        a %= p
        if a == 0:
            return 0
        if pow(a, (p - 1) // 2, p) != 1:
            return None
        if p % 4 == 3:
            return pow(a, (p + 1) // 4, p)
        s, t = 0, p - 1
        while t % 2 == 0:
            s += 1
            t //= 2
        z = 2
        while pow(z, (p - 1) // 2, p) != p - 1:
            z += 1
        m, c, tt, rr = s, pow(z, t, p), pow(a, t, p), pow(a, (t + 1) // 2, p)
        while tt != 1:
            i, t2 = 0, tt
            while t2 != 1:
                t2 = t2 * t2 % p
                i += 1
            b = pow(c, 1 << (m - i - 1), p)
            m, c = i, b * b % p
            tt, rr = tt * c % p, rr * b % p
        return rr


class Fq2Ops:
    """Fq2 = Fq[u]/(u^2 - beta); elements are (c0, c1) tuples."""

    def __init__(self, p, beta):
        self.p = p
        self.beta = beta % p
        self.zero = (0, 0)
        self.one = (1, 0)

    def add(self, a, b): return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)
    def sub(self, a, b): return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)
    def neg(self, a): return ((-a[0]) % self.p, (-a[1]) % self.p)

    def mul(self, a, b):
        p = self.p
        return ((a[0] * b[0] + self.beta * a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

    def inv(self, a):
        p = self.p
        n = (a[0] * a[0] - self.beta * a[1] * a[1]) % p
        ni = pow(n, -1, p)
        return (a[0] * ni % p, (-a[1]) * ni % p)

    def is_zero(self, a): return a[0] % self.p == 0 and a[1] % self.p == 0
    def eq(self, a, b): return self.is_zero(self.sub(a, b))
    def from_int(self, a): return (a % self.p, 0)

    def pow(self, a, e):
        res = self.one
        base = a
        while e:
            if e & 1:
                res = self.mul(res, base)
            base = self.mul(base, base)
            e >>= 1
        return res

    def sqrt(self, a):
        """Generic Tonelli-Shanks in Fq2 (group order q^2 - 1)."""
        if self.is_zero(a):
            return self.zero
        p = self.p
        order = p * p - 1
        if self.pow(a, order // 2) != self.one:
            return None
        s, t = 0, order
        while t % 2 == 0:
            s += 1
            t //= 2
        # deterministic search for a non-residue
        k = 1
        while True:
            z = (k, 1)
            if self.pow(z, order // 2) != self.one:
                break
            k += 1
        m, c, tt, rr = s, self.pow(z, t), self.pow(a, t), self.pow(a, (t + 1) // 2)
        while tt != self.one:
            i, t2 = 0, tt
            while t2 != self.one:
                t2 = self.mul(t2, t2)
                i += 1
            b = self.pow(c, 1 << (m - i - 1))
            m, c = i, self.mul(b, b)
            tt, rr = self.mul(tt, c), self.mul(rr, b)
        return rr


# --------------------------------------------------------------------------------------------
# Short-Weierstrass group law (a = 0) in affine coordinates over a FieldOps; None = identity
# --------------------------------------------------------------------------------------------


class Group:
    def __init__(self, F, b):
        self.F = F
        self.b = b

    def on_curve(self, P):
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.eq(F.mul(y, y), F.add(F.mul(F.mul(x, x), x), self.b))

    def neg(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))

    def add(self, P, Q):
        F = self.F
        if P is None:
            return Q
        if Q is None:
            return P
        if F.eq(P[0], Q[0]):
            if F.eq(P[1], Q[1]):
                return self.dbl(P)
            return None
        lam = F.mul(F.sub(Q[1], P[1]), F.inv(F.sub(Q[0], P[0])))
        x3 = F.sub(F.sub(F.mul(lam, lam), P[0]), Q[0])
        y3 = F.sub(F.mul(lam, F.sub(P[0], x3)), P[1])
        return (x3, y3)

    def dbl(self, P):
        F = self.F
        if P is None or F.is_zero(P[1]):
            return None
        x, y = P
        xx = F.mul(x, x)
        lam = F.mul(F.add(F.add(xx, xx), xx), F.inv(F.add(y, y)))
        x3 = F.sub(F.sub(F.mul(lam, lam), x), x)
        y3 = F.sub(F.mul(lam, F.sub(x, x3)), y)
        return (x3, y3)

    # Jacobian internals for fast scalar multiplication (no inversions in the loop)
    def _jdbl(self, P):
        F = self.F
        X, Y, Z = P
        if F.is_zero(Z):
            return P
        A = F.mul(X, X)
        B = F.mul(Y, Y)
        C = F.mul(B, B)
        t = F.add(X, B)
        D = F.sub(F.sub(F.mul(t, t), A), C)
        D = F.add(D, D)
        E = F.add(F.add(A, A), A)
        Fv = F.mul(E, E)
        X3 = F.sub(Fv, F.add(D, D))
        C8 = F.add(C, C); C8 = F.add(C8, C8); C8 = F.add(C8, C8)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
        Z3 = F.mul(F.add(Y, Y), Z)
        return (X3, Y3, Z3)

    def _jadd(self, P, Q):
        F = self.F
        if F.is_zero(P[2]):
            return Q
        if F.is_zero(Q[2]):
            return P
        X1, Y1, Z1 = P
        X2, Y2, Z2 = Q
        Z1Z1 = F.mul(Z1, Z1)
        Z2Z2 = F.mul(Z2, Z2)
        U1 = F.mul(X1, Z2Z2)
        U2 = F.mul(X2, Z1Z1)
        S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
        S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
        if F.eq(U1, U2):
            if F.eq(S1, S2):
                return self._jdbl(P)
            return (F.one, F.one, F.zero)
        H = F.sub(U2, U1)
        R = F.sub(S2, S1)
        HH = F.mul(H, H)
        HHH = F.mul(H, HH)
        V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.mul(R, R), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(R, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def to_jac(self, P):
        F = self.F
        return (F.one, F.one, F.zero) if P is None else (P[0], P[1], F.one)

    def from_jac(self, P):
        F = self.F
        if F.is_zero(P[2]):
            return None
        zi = F.inv(P[2])
        zi2 = F.mul(zi, zi)
        return (F.mul(P[0], zi2), F.mul(P[1], F.mul(zi2, zi)))

    def mul(self, P, k: int):
        if P is None or k == 0:
            return None
        if k < 0:
            return self.mul(self.neg(P), -k)
        acc = (self.F.one, self.F.one, self.F.zero)
        base = self.to_jac(P)
        for bit in bin(k)[2:]:
            acc = self._jdbl(acc)
            if bit == "1":
                acc = self._jadd(acc, base)
        return self.from_jac(acc)

    def msm_naive(self, bases, scalars):
        """sum scalars[i]*bases[i], truncated to the shorter length like ark-ec msm_bigint (section 2a)."""
        n = min(len(bases), len(scalars))
        acc = (self.F.one, self.F.one, self.F.zero)
        for i in range(n):
            if scalars[i] == 0 or bases[i] is None:
                continue
            acc = self._jadd(acc, self.to_jac(self.mul(bases[i], scalars[i])))
        return self.from_jac(acc)


# --------------------------------------------------------------------------------------------
# Per-curve context: groups, twist, subgroup generators, Fq12 for the pairing
# --------------------------------------------------------------------------------------------


class CurveCtx:
    def __init__(self, c: Curve):
        self.c = c
        self.Fq = FqOps(c.q)
        self.Fq2 = Fq2Ops(c.q, c.beta)
        self.G1 = Group(self.Fq, c.b % c.q)
        xi = (c.xi[0] % c.q, c.xi[1] % c.q)
        if c.twist == "M":
            b2 = self.Fq2.mul((c.b, 0), xi)
        else:
            b2 = self.Fq2.mul((c.b, 0), self.Fq2.inv(xi))
        self.b2 = b2
        self.G2 = Group(self.Fq2, b2)
        # trace of Frobenius
        if c.family == "bls12":
            self.t = c.x + 1
        else:
            self.t = 6 * c.x * c.x + 1
        self.n1 = c.q + 1 - self.t
        assert self.n1 % c.r == 0
        self.h1 = self.n1 // c.r
        # order of the correct sextic twist over Fq2
        t2 = self.t * self.t - 2 * c.q
        f2sq = (4 * c.q * c.q - t2 * t2) // 3
        f2 = _isqrt(f2sq)
        assert f2 * f2 == f2sq
        cands = [c.q * c.q + 1 - (t2 + 3 * f2) // 2, c.q * c.q + 1 - (t2 - 3 * f2) // 2,
                 c.q * c.q + 1 - (-t2 + 3 * f2) // 2, c.q * c.q + 1 - (-t2 - 3 * f2) // 2]
        self._n2_cands = [n for n in cands if n % c.r == 0]
        self._g1 = None
        self._g2 = None
        # Fq12 as Fq[w]/(w^12 - 2a w^6 + a^2 - beta b^2)
        a, bb = xi
        self.f12_c6 = (2 * a) % c.q  # w^12 = c6*w^6 - c0
        self.f12_c0 = (a * a - c.beta * bb * bb) % c.q
        self.xi = xi

    # -- deterministic subgroup generators (the reference samples random ones, generator.rs:26-32) --
    def g1_gen(self):
        if self._g1 is None:
            self._g1 = self._find_point(self.G1, self.Fq, [self.h1], b"g1")
        return self._g1

    def g2_gen(self):
        if self._g2 is None:
            self._g2 = self._find_point(self.G2, self.Fq2, [n // self.c.r for n in self._n2_cands], b"g2")
        return self._g2

    def _find_point(self, G, F, cofactors, tag):
        ctr = 0
        while True:
            seed = int.from_bytes(hashlib.sha256(tag + self.c.name.encode() + bytes([ctr])).digest(), "big")
            ctr += 1
            x = F.from_int(seed) if F is self.Fq else (seed % self.c.q, (seed >> 7) % self.c.q)
            rhs = F.add(F.mul(F.mul(x, x), x), G.b)
            y = F.sqrt(rhs)
            if y is None:
                continue
            P = (x, y)
            assert G.on_curve(P)
            for h in cofactors:
                Q = G.mul(P, h)
                if Q is not None and G.mul(Q, self.c.r) is None:
                    return Q

    # -- Fq12 polynomial arithmetic (coefficient lists of length 12 over Fq) --
    def f12_mul(self, a, b):
        q = self.c.q
        t = [0] * 23
        for i, ai in enumerate(a):
            if ai == 0:
                continue
            for j, bj in enumerate(b):
                t[i + j] += ai * bj
        for k in range(22, 11, -1):
            v = t[k] % q
            if v:
                t[k - 6] += v * self.f12_c6
                t[k - 12] -= v * self.f12_c0
        return [v % q for v in t[:12]]

    def f12_pow(self, a, e):
        res = [1] + [0] * 11
        base = a
        while e:
            if e & 1:
                res = self.f12_mul(res, base)
            base = self.f12_mul(base, base)
            e >>= 1
        return res

    def fq2_to_f12(self, a, shift=0):
        """Embed c0 + c1*u, u = (w^6 - xi0)/xi1, then multiply by w^shift (shift may be negative)."""
        q = self.c.q
        xa, xb = self.xi
        xbi = inv(xb, q)
        out = [0] * 12
        out[0] = (a[0] - a[1] * xa * xbi) % q
        out[6] = (a[1] * xbi) % q
        if shift > 0:
            wk = [0] * 12
            wk[shift] = 1
            out = self.f12_mul(out, wk)
        elif shift < 0:
            # w^-1 = w^11 * (w^12)^-1 ; compute via generic inverse of w^|shift| by exponentiation
            wk = [0] * 12
            wk[-shift] = 1
            out = self.f12_mul(out, self.f12_inv(wk))
        return out

    def f12_inv(self, a):
        # a^(q^12 - 2)
        return self.f12_pow(a, self.c.q ** 12 - 2)

    def untwist(self, Q):
        """E'(Fq2) -> E(Fq12): D-type (x w^2, y w^3); M-type (x / w^2, y / w^3)."""
        if self.c.twist == "D":
            return (self.fq2_to_f12(Q[0], 2), self.fq2_to_f12(Q[1], 3))
        return (self.fq2_to_f12(Q[0], -2), self.fq2_to_f12(Q[1], -3))

    def miller_tate(self, P, Q):
        """f_{r,P}(Q) with P in G1 (affine over Fq) and Q in G2; reduced Tate pairing after final exp.
        Any non-degenerate bilinear pairing validates the Groth16 equation (verifier.rs:44-65)."""
        one = [1] + [0] * 11
        if P is None or Q is None:
            return one
        q = self.c.q
        xq, yq = self.untwist(Q)
        f = one
        T = P
        G1 = self.G1
        bits = bin(self.c.r)[3:]

        def line(T, R):
            # line through T and R (or tangent) evaluated at (xq, yq); vertical lines are killed by the final exp
            if T[0] == R[0] and T[1] == R[1]:
                lam = 3 * T[0] * T[0] * inv(2 * T[1], q) % q
            elif T[0] == R[0]:
                # vertical: xq - xT  (lies in a proper subfield only for even powers; keep it, harmless)
                out = list(xq)
                out[0] = (out[0] - T[0]) % q
                return out
            else:
                lam = (R[1] - T[1]) * inv(R[0] - T[0], q) % q
            # yq - yT - lam (xq - xT)
            out = [(yq[i] - lam * xq[i]) % q for i in range(12)]
            out[0] = (out[0] - T[1] + lam * T[0]) % q
            return out

        for bit in bits:
            f = self.f12_mul(self.f12_mul(f, f), line(T, T))
            T = G1.dbl(T)
            if bit == "1":
                if T is None:
                    T = P
                else:
                    f = self.f12_mul(f, line(T, P))
                    T = G1.add(T, P)
        return f

    def final_exp(self, f):
        return self.f12_pow(f, (self.c.q ** 12 - 1) // self.c.r)

    def pairing_product_is_one(self, pairs):
        f = [1] + [0] * 11
        for P, Q in pairs:
            f = self.f12_mul(f, self.miller_tate(P, Q))
        return self.final_exp(f) == [1] + [0] * 11


def _isqrt(n):
    import math
    return math.isqrt(n)


_CTX = {}


def ctx(curve) -> CurveCtx:
    c = curve if isinstance(curve, Curve) else CURVES[curve]
    if c.name not in _CTX:
        _CTX[c.name] = CurveCtx(c)
    return _CTX[c.name]


# --------------------------------------------------------------------------------------------
# Radix-2 evaluation domain (ark-poly Radix2EvaluationDomain behaviour, section 2a)
# --------------------------------------------------------------------------------------------


class Domain:
    def __init__(self, c: Curve, min_size: int):
        n = 1
        log_n = 0
        while n < min_size:
            n <<= 1
            log_n += 1
        if log_n > c.two_adicity:
            raise ValueError("PolynomialDegreeTooLarge")  # r1cs_to_qap.rs:179
        self.c, self.r, self.n, self.log_n = c, c.r, n, log_n
        two_adic_root = pow(c.fr_gen, (c.r - 1) >> c.two_adicity, c.r)
        self.omega = pow(two_adic_root, 1 << (c.two_adicity - log_n), c.r)
        self.omega_inv = inv(self.omega, c.r)
        self.n_inv = inv(n, c.r)

    def _fft_core(self, a, root):
        """In-order radix-2 transform: out[k] = sum_j a[j] root^(jk)."""
        r, n = self.r, self.n
        a = list(a) + [0] * (n - len(a))
        # bit reversal then DIT
        j = 0
        for i in range(1, n):
            bit = n >> 1
            while j & bit:
                j ^= bit
                bit >>= 1
            j |= bit
            if i < j:
                a[i], a[j] = a[j], a[i]
        m = 2
        while m <= n:
            wm = pow(root, n // m, r)
            half = m // 2
            for k in range(0, n, m):
                w = 1
                for t in range(half):
                    u = a[k + t]
                    v = a[k + t + half] * w % r
                    a[k + t] = (u + v) % r
                    a[k + t + half] = (u - v) % r
                    w = w * wm % r
            m <<= 1
        return a

    def fft(self, coeffs, offset=1):
        r = self.r
        a = list(coeffs) + [0] * (self.n - len(coeffs))
        if offset != 1:
            g = 1
            for i in range(self.n):
                a[i] = a[i] * g % r
                g = g * offset % r
        return self._fft_core(a, self.omega)

    def ifft(self, evals, offset=1):
        r = self.r
        a = self._fft_core(evals, self.omega_inv)
        if offset == 1:
            return [x * self.n_inv % r for x in a]
        oi = inv(offset, r)
        g = self.n_inv
        out = []
        for x in a:
            out.append(x * g % r)
            g = g * oi % r
        return out

    def dft_naive(self, coeffs, offset=1):
        r = self.r
        out = []
        for k in range(self.n):
            x = offset * pow(self.omega, k, r) % r
            acc = 0
            for cf in reversed(coeffs):
                acc = (acc * x + cf) % r
            out.append(acc)
        return out

    def vanishing(self, tau):
        return (pow(tau, self.n, self.r) - 1) % self.r

    def lagrange_coeffs(self, tau):
        """evaluate_all_lagrange_coefficients(tau) for tau outside the domain: L_i(tau) = Z(tau) w^i / (n (tau - w^i))."""
        r, n = self.r, self.n
        z = self.vanishing(tau)
        assert z != 0
        dens = []
        w = 1
        for i in range(n):
            dens.append((tau - w) % r)
            w = w * self.omega % r
        invs = batch_inv(dens, r)
        out = []
        w = 1
        zn = z * self.n_inv % r
        for i in range(n):
            out.append(zn * w % r * invs[i] % r)
            w = w * self.omega % r
        return out


def batch_inv(xs, p):
    pref = []
    acc = 1
    for x in xs:
        pref.append(acc)
        acc = acc * x % p
    ai = inv(acc, p)
    out = [0] * len(xs)
    for i in range(len(xs) - 1, -1, -1):
        out[i] = ai * pref[i] % p
        ai = ai * xs[i] % p
    return out


# --------------------------------------------------------------------------------------------
# R1CS as ConstraintMatrices (ark-relations to_matrices(): rows of (coeff, column); column <
# num_instance => instance variable, else witness index + num_instance).  SURVEY section 7 step 1.
# --------------------------------------------------------------------------------------------


@dataclass
class R1CS:
    curve: Curve
    num_instance: int  # includes the constant One at column 0
    num_witness: int
    a: List[List[Tuple[int, int]]]
    b: List[List[Tuple[int, int]]]
    c: List[List[Tuple[int, int]]]
    assignment: Optional[List[int]] = None  # instance || witness, canonical ints

    @property
    def num_constraints(self):
        return len(self.a)

    def is_satisfied(self):
        r = self.curve.r
        z = self.assignment
        for ra, rb, rc in zip(self.a, self.b, self.c):
            ea = sum(cf * z[i] for cf, i in ra) % r
            eb = sum(cf * z[i] for cf, i in rb) % r
            ec = sum(cf * z[i] for cf, i in rc) % r
            if ea * eb % r != ec:
                return False
        return True


class Rng:
    """splitmix64 stream; scalars by rejection sampling < r."""

    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def u64(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def below(self, n):
        return self.u64() % n

    def fr(self, r):
        bits = r.bit_length()
        while True:
            v = 0
            for i in range(4):
                v |= self.u64() << (64 * i)
            v &= (1 << bits) - 1
            if v < r:
                return v


def silly_circuit(curve: Curve, a: int, b: int) -> R1CS:
    """MySillyCircuit (src/test.rs:14-43): witness a,b; input c = a*b; six copies of a*b=c."""
    r = curve.r
    row_a, row_b, row_c = [(1, 2)], [(1, 3)], [(1, 1)]
    return R1CS(curve, 2, 2, [list(row_a) for _ in range(6)], [list(row_b) for _ in range(6)],
                [list(row_c) for _ in range(6)], [1, a * b % r, a % r, b % r])


MIMC_ROUNDS = 322


def mimc_hash(curve, xl, xr, constants):
    r = curve.r
    for cst in constants:
        t = (xl + cst) % r
        t = (t * t % r * t + xr) % r
        xr, xl = xl, t
    return xl


def mimc_circuit(curve: Curve, xl: int, xr: int, constants: Sequence[int]) -> R1CS:
    """MiMCDemo (tests/mimc.rs:65-143).  Witness order: xl, xr, then per round tmp, new_xl; the last
    new_xl is the single public input (instance column 1)."""
    r = curve.r
    ninst = 2
    A, B, C = [], [], []
    wit = [xl % r, xr % r]
    xl_col, xr_col = ninst + 0, ninst + 1
    xl_v, xr_v = xl % r, xr % r
    image = None
    for i, cst in enumerate(constants):
        cst %= r
        tmp_v = (xl_v + cst) ** 2 % r
        wit.append(tmp_v)
        tmp_col = ninst + len(wit) - 1
        lin = [(1, xl_col)] + ([(cst, 0)] if cst else [])
        A.append(list(lin)); B.append(list(lin)); C.append([(1, tmp_col)])
        new_v = ((xl_v + cst) * tmp_v + xr_v) % r
        if i == len(constants) - 1:
            new_col = 1
            image = new_v
        else:
            wit.append(new_v)
            new_col = ninst + len(wit) - 1
        A.append([(1, tmp_col)]); B.append(list(lin)); C.append([(1, new_col), (r - 1, xr_col)])
        xr_col, xr_v = xl_col, xl_v
        xl_col, xl_v = new_col, new_v
    return R1CS(curve, ninst, len(wit), A, B, C, [1, image] + wit)


def dummy_circuit(curve: Curve, a: int, b: int, num_variables: int, num_constraints: int) -> R1CS:
    """DummyCircuit (benches/bench.rs:41-64): witness a, b, then num_variables-3 copies of a;
    input c = a*b; num_constraints-1 copies of a*b=c and one empty constraint."""
    r = curve.r
    wit = [a % r, b % r] + [a % r] * (num_variables - 3)
    A = [[(1, 2)] for _ in range(num_constraints - 1)] + [[]]
    B = [[(1, 3)] for _ in range(num_constraints - 1)] + [[]]
    C = [[(1, 1)] for _ in range(num_constraints - 1)] + [[]]
    return R1CS(curve, 2, len(wit), A, B, C, [1, a * b % r] + wit)


def synthetic_circuit(curve: Curve, num_constraints: int, seed: int, num_inputs: int = 1) -> R1CS:
    """Non-degenerate synthetic R1CS: constraint i is (z_p + k_i) * z_q = z_new with p, q
    uniform over earlier variables, k_i uniform in Fr; the last `num_inputs` products are public inputs."""
    r = curve.r
    rng = Rng(seed)
    ninst = 1 + num_inputs
    nwit = 2 + num_constraints - num_inputs
    vals = {}
    cols = []  # column of each variable in creation order
    # two seed witnesses
    wit_vals = [rng.fr(r), rng.fr(r)]
    cols = [ninst + 0, ninst + 1]
    allv = list(wit_vals)
    inst_vals = []
    A, B, C = [], [], []
    for i in range(num_constraints):
        p = rng.below(len(cols))
        qq = rng.below(len(cols))
        k = rng.fr(r)
        v = (allv[p] + k) * allv[qq] % r
        if i >= num_constraints - num_inputs:
            col = 1 + len(inst_vals)
            inst_vals.append(v)
        else:
            col = ninst + len(wit_vals)
            wit_vals.append(v)
        A.append([(1, cols[p]), (k, 0)])
        B.append([(1, cols[qq])])
        C.append([(1, col)])
        cols.append(col)
        allv.append(v)
    assert len(wit_vals) == nwit
    return R1CS(curve, ninst, nwit, A, B, C, [1] + inst_vals + wit_vals)


# --------------------------------------------------------------------------------------------
# R1CS -> QAP (r1cs_to_qap.rs)
# --------------------------------------------------------------------------------------------


def evaluate_constraint(terms, z, r):
    """r1cs_to_qap.rs:28-67."""
    return sum(cf * z[i] for cf, i in terms) % r


def abc_evals(cs: R1CS):
    """The three domain-sized vectors fed to the iFFTs (r1cs_to_qap.rs:183-199,213-218)."""
    r = cs.curve.r
    dom = Domain(cs.curve, cs.num_constraints + cs.num_instance)
    z = cs.assignment
    a = [evaluate_constraint(row, z, r) for row in cs.a] + [0] * (dom.n - cs.num_constraints)
    b = [evaluate_constraint(row, z, r) for row in cs.b] + [0] * (dom.n - cs.num_constraints)
    c = [evaluate_constraint(row, z, r) for row in cs.c] + [0] * (dom.n - cs.num_constraints)
    for i in range(cs.num_instance):
        a[cs.num_constraints + i] = z[i]
    return dom, a, b, c


def witness_map_from_evals(dom: Domain, a, b, c):
    """r1cs_to_qap.rs:201-234: 3 iFFT, 3 coset FFT, (ab - c)/Z on the coset, coset iFFT."""
    r = dom.r
    g = dom.c.fr_gen
    a = dom.fft(dom.ifft(a), offset=g)
    b = dom.fft(dom.ifft(b), offset=g)
    c = dom.fft(dom.ifft(c), offset=g)
    zinv = inv(dom.vanishing(g), r)
    ab = [((x * y - w) % r) * zinv % r for x, y, w in zip(a, b, c)]
    return dom.ifft(ab, offset=g)


def witness_map(cs: R1CS):
    dom, a, b, c = abc_evals(cs)
    return witness_map_from_evals(dom, a, b, c)


def instance_map_with_evaluation(cs: R1CS, t: int):
    """r1cs_to_qap.rs:128-170."""
    r = cs.curve.r
    dom = Domain(cs.curve, cs.num_constraints + cs.num_instance)
    zt = dom.vanishing(t)
    u = dom.lagrange_coeffs(t)
    nvar = (cs.num_instance - 1) + cs.num_witness
    a = [0] * (nvar + 1)
    b = [0] * (nvar + 1)
    c = [0] * (nvar + 1)
    nc = cs.num_constraints
    for i in range(cs.num_instance):
        a[i] = u[nc + i]
    for i in range(nc):
        ui = u[i]
        for cf, idx in cs.a[i]:
            a[idx] = (a[idx] + ui * cf) % r
        for cf, idx in cs.b[i]:
            b[idx] = (b[idx] + ui * cf) % r
        for cf, idx in cs.c[i]:
            c[idx] = (c[idx] + ui * cf) % r
    return a, b, c, zt, nvar, dom.n


# --------------------------------------------------------------------------------------------
# Keys / proof and the three protocol functions
# --------------------------------------------------------------------------------------------


@dataclass
class VerifyingKey:
    alpha_g1: object
    beta_g2: object
    gamma_g2: object
    delta_g2: object
    gamma_abc_g1: list


@dataclass
class ProvingKey:
    """data_structures.rs:126-143."""
    curve: Curve
    vk: VerifyingKey
    beta_g1: object
    delta_g1: object
    a_query: list
    b_g1_query: list
    b_g2_query: list
    h_query: list
    l_query: list
    toxic: Optional[dict] = None  # oracle-only: the toxic waste, for "proof in the exponent" checks


@dataclass
class Proof:
    a: object
    b: object
    c: object


def generate_parameters(cs: R1CS, alpha, beta, gamma, delta, tau, g1=None, g2=None,
                        scalars_only=False):
    """generator.rs:47-208 with explicit toxic waste (the reference samples it, generator.rs:19-44).
    With scalars_only=True returns the exponent vectors (used to drive the GPU fixed-base path)."""
    cv = cs.curve
    r = cv.r
    cx = ctx(cv)
    g1 = g1 or cx.g1_gen()
    g2 = g2 or cx.g2_gen()
    a, b, c, zt, nvar, m_raw = instance_map_with_evaluation(cs, tau)
    gi, di = inv(gamma, r), inv(delta, r)
    ni = cs.num_instance
    gamma_abc = [(beta * a[i] + alpha * b[i] + c[i]) * gi % r for i in range(ni)]
    l = [(beta * a[i] + alpha * b[i] + c[i]) * di % r for i in range(ni, nvar + 1)]
    # h_query_scalars (r1cs_to_qap.rs:237-247), max_power = m_raw - 1
    hs = []
    tp = zt * di % r
    for i in range(m_raw - 1):
        hs.append(tp)
        tp = tp * tau % r
    exps = dict(a=a, b=b, l=l, h=hs, gamma_abc=gamma_abc, alpha=alpha, beta=beta, gamma=gamma,
                delta=delta, tau=tau, zt=zt)
    if scalars_only:
        return exps
    G1, G2 = cx.G1, cx.G2
    vk = VerifyingKey(G1.mul(g1, alpha), G2.mul(g2, beta), G2.mul(g2, gamma), G2.mul(g2, delta),
                      [G1.mul(g1, s) for s in gamma_abc])
    pk = ProvingKey(cv, vk, G1.mul(g1, beta), G1.mul(g1, delta),
                    [G1.mul(g1, s) for s in a], [G1.mul(g1, s) for s in b], [G2.mul(g2, s) for s in b],
                    [G1.mul(g1, s) for s in hs], [G1.mul(g1, s) for s in l],
                    toxic=dict(exps, g1=g1, g2=g2))
    return pk


def create_proof_with_assignment(pk: ProvingKey, r_: int, s_: int, h, input_assignment, aux_assignment,
                                 msm1=None, msm2=None):
    """prover.rs:54-132.  msm1/msm2 let tests substitute another MSM implementation."""
    cv = pk.curve
    cx = ctx(cv)
    G1, G2 = cx.G1, cx.G2
    msm1 = msm1 or G1.msm_naive
    msm2 = msm2 or G2.msm_naive
    h_acc = msm1(pk.h_query, h)
    l_aux_acc = msm1(pk.l_query, aux_assignment)
    r_s_delta_g1 = G1.mul(pk.delta_g1, r_ * s_ % cv.r)
    assignment = list(input_assignment) + list(aux_assignment)

    def calc(G, msm, initial, query, vk_param):  # prover.rs:252-270
        acc = msm(query[1:], assignment)
        res = G.add(initial, query[0])
        res = G.add(res, acc)
        return G.add(res, vk_param)

    g_a = calc(G1, msm1, G1.mul(pk.delta_g1, r_), pk.a_query, pk.vk.alpha_g1)
    s_g_a = G1.mul(g_a, s_)
    if r_ % cv.r != 0:
        g1_b = calc(G1, msm1, G1.mul(pk.delta_g1, s_), pk.b_g1_query, pk.beta_g1)
    else:
        g1_b = None
    g2_b = calc(G2, msm2, G2.mul(pk.vk.delta_g2, s_), pk.b_g2_query, pk.vk.beta_g2)
    r_g1_b = G1.mul(g1_b, r_)
    g_c = G1.add(s_g_a, r_g1_b)
    g_c = G1.add(g_c, G1.neg(r_s_delta_g1))
    g_c = G1.add(g_c, l_aux_acc)
    g_c = G1.add(g_c, h_acc)
    return Proof(g_a, g2_b, g_c)


def create_proof(pk: ProvingKey, cs: R1CS, r_: int, s_: int, **kw):
    """create_proof_with_reduction_and_matrices (prover.rs:26-51)."""
    h = witness_map(cs)
    z = cs.assignment
    return create_proof_with_assignment(pk, r_, s_, h, z[1:cs.num_instance], z[cs.num_instance:], **kw)


def verify_proof(vk: VerifyingKey, curve: Curve, proof: Proof, public_inputs) -> bool:
    """verifier.rs:25-76: e(A,B) == e(alpha,beta) e(g_ic,gamma) e(C,delta), as a pairing product == 1."""
    cx = ctx(curve)
    G1 = cx.G1
    if len(public_inputs) + 1 != len(vk.gamma_abc_g1):
        raise ValueError("MalformedVerifyingKey")  # verifier.rs:29-31
    g_ic = vk.gamma_abc_g1[0]
    for x, base in zip(public_inputs, vk.gamma_abc_g1[1:]):
        g_ic = G1.add(g_ic, G1.mul(base, x % curve.r))
    return cx.pairing_product_is_one([
        (proof.a, proof.b),
        (G1.neg(vk.alpha_g1), vk.beta_g2),
        (G1.neg(g_ic), vk.gamma_g2),
        (G1.neg(proof.c), vk.delta_g2),
    ])


def proof_in_the_exponent(pk: ProvingKey, cs: R1CS, r_: int, s_: int, h=None) -> Proof:
    """SURVEY section 8c check 5: with the toxic waste known, A, B, C are three scalar multiplications."""
    cv = pk.curve
    rr = cv.r
    tx = pk.toxic
    cx = ctx(cv)
    z = cs.assignment
    a_t = sum(zi * ai for zi, ai in zip(z, tx["a"])) % rr
    b_t = sum(zi * bi for zi, bi in zip(z, tx["b"])) % rr
    if h is None:
        h = witness_map(cs)
    A = (tx["alpha"] + a_t + r_ * tx["delta"]) % rr
    B = (tx["beta"] + b_t + s_ * tx["delta"]) % rr
    l_t = sum(zi * li for zi, li in zip(z[cs.num_instance:], tx["l"])) % rr
    h_t = sum(hi * si for hi, si in zip(h, tx["h"])) % rr  # truncated to n-1 like msm_bigint
    Cx = (l_t + h_t + s_ * A + r_ * B - r_ * s_ % rr * tx["delta"]) % rr
    return Proof(cx.G1.mul(tx["g1"], A), cx.G2.mul(tx["g2"], B), cx.G1.mul(tx["g1"], Cx))
