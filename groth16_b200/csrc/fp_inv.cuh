// fp_inv.cuh -- modular inversion by Bernstein-Yang "safegcd" division steps (signed 30-bit limbs, 30 steps per
// 2x2 transition matrix), for the places where a Fermat inversion (~1.5 * bits multiplications, ~190 k instructions
// for a 381-bit field) sits on a latency-critical path: the one-inversion-per-lane combine kernel of msm_ba.cuh.
// ~37 (384-bit) / 25 (256-bit) matrix steps of ~500 straight-line instructions, no data-dependent branches.
//
// Step count: 30 * ceil(((49 bits + 57) / 17) / 30), the proven bound for the plain delta = 1 rule; the
// "half-delta" rule used here (zeta = -(delta + 1/2), start -1) needs fewer, and surplus steps are harmless
// (g stays 0, f stays +-1).  Checked against Fp::inv (Fermat) on the CPU for all six fields: tests/host/inv_check.cu.
#pragma once
#include "fp.cuh"

namespace g16 {

template <int L>
struct Signed30 {
  int32_t v[L];
};

// 30 division steps on the low words; returns the new zeta and t = (u, v; q, r) with t * (f, g) = 2^30 * (f', g')
G16_HD int32_t safegcd_divsteps_30(int32_t zeta, uint32_t f0, uint32_t g0, int32_t t[4]) {
  uint32_t u = 1, v = 0, q = 0, r = 1;
  uint32_t f = f0, g = g0;
#pragma unroll
  for (int i = 0; i < 30; i++) {
    uint32_t c1 = (uint32_t)(zeta >> 31);    // all ones iff zeta < 0
    const uint32_t c2 = 0u - (g & 1u);       // all ones iff g odd
    const uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;   // conditionally negated f, u, v
    g += x & c2;
    q += y & c2;
    r += z & c2;
    c1 &= c2;                                 // swap iff zeta < 0 and g odd
    zeta = (int32_t)((uint32_t)zeta ^ c1) - 1;
    f += g & c1;
    u += q & c1;
    v += r & c1;
    g >>= 1;
    u <<= 1;
    v <<= 1;
  }
  t[0] = (int32_t)u; t[1] = (int32_t)v; t[2] = (int32_t)q; t[3] = (int32_t)r;
  return zeta;
}

// (d, e) <- t * (d, e) / 2^30 mod m, entries kept in (-2m, m)
template <int L>
G16_HD void safegcd_update_de(Signed30<L>& d, Signed30<L>& e, const int32_t t[4], const Signed30<L>& m, uint32_t m_inv30) {
  constexpr int32_t M30 = (1 << 30) - 1;
  const int32_t u = t[0], v = t[1], q = t[2], r = t[3];
  const int32_t sd = d.v[L - 1] >> 31, se = e.v[L - 1] >> 31;
  int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
  int32_t di = d.v[0], ei = e.v[0];
  int64_t cd = (int64_t)u * di + (int64_t)v * ei, ce = (int64_t)q * di + (int64_t)r * ei;
  md -= (int32_t)((m_inv30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
  me -= (int32_t)((m_inv30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
  cd += (int64_t)m.v[0] * md;
  ce += (int64_t)m.v[0] * me;
  cd >>= 30;   // the low 30 bits are zero by construction
  ce >>= 30;
#pragma unroll
  for (int i = 1; i < L; i++) {
    di = d.v[i];
    ei = e.v[i];
    cd += (int64_t)u * di + (int64_t)v * ei;
    ce += (int64_t)q * di + (int64_t)r * ei;
    cd += (int64_t)m.v[i] * md;
    ce += (int64_t)m.v[i] * me;
    d.v[i - 1] = (int32_t)cd & M30;
    cd >>= 30;
    e.v[i - 1] = (int32_t)ce & M30;
    ce >>= 30;
  }
  d.v[L - 1] = (int32_t)cd;
  e.v[L - 1] = (int32_t)ce;
}

// (f, g) <- t * (f, g) / 2^30 (exact)
template <int L>
G16_HD void safegcd_update_fg(Signed30<L>& f, Signed30<L>& g, const int32_t t[4]) {
  constexpr int32_t M30 = (1 << 30) - 1;
  const int32_t u = t[0], v = t[1], q = t[2], r = t[3];
  int32_t fi = f.v[0], gi = g.v[0];
  int64_t cf = (int64_t)u * fi + (int64_t)v * gi, cg = (int64_t)q * fi + (int64_t)r * gi;
  cf >>= 30;
  cg >>= 30;
#pragma unroll
  for (int i = 1; i < L; i++) {
    fi = f.v[i];
    gi = g.v[i];
    cf += (int64_t)u * fi + (int64_t)v * gi;
    cg += (int64_t)q * fi + (int64_t)r * gi;
    f.v[i - 1] = (int32_t)cf & M30;
    cf >>= 30;
    g.v[i - 1] = (int32_t)cg & M30;
    cg >>= 30;
  }
  f.v[L - 1] = (int32_t)cf;
  g.v[L - 1] = (int32_t)cg;
}

// r in (-2m, m) -> [0, m), negated first when sign < 0
template <int L>
G16_HD void safegcd_normalize(Signed30<L>& r, int32_t sign, const Signed30<L>& m) {
  constexpr int32_t M30 = (1 << 30) - 1;
  int32_t cond_add = r.v[L - 1] >> 31;
  const int32_t cond_negate = sign >> 31;
#pragma unroll
  for (int i = 0; i < L; i++) {
    r.v[i] += m.v[i] & cond_add;
    r.v[i] = (r.v[i] ^ cond_negate) - cond_negate;
  }
#pragma unroll
  for (int i = 0; i < L - 1; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= M30; }
  cond_add = r.v[L - 1] >> 31;
#pragma unroll
  for (int i = 0; i < L; i++) r.v[i] += m.v[i] & cond_add;
#pragma unroll
  for (int i = 0; i < L - 1; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= M30; }
}

// a^-1 in the Montgomery domain (a = xR  ->  x^-1 R); inverse of zero is zero -- same contract as Fp::inv
template <class P>
G16_HD Fp<P> fp_inv_safegcd(const Fp<P>& a) {
  constexpr int N = P::N;
  constexpr int L = (32 * N + 29) / 30;
  constexpr int ITER = ((49 * P::BITS + 57) / 17 + 29) / 30;
  constexpr uint32_t M30 = (1u << 30) - 1;
  if (a.is_zero()) return a;
  Signed30<L> m, d, e, f, g;
#pragma unroll
  for (int i = 0; i < L; i++) {
    const int bit = 30 * i, k = bit >> 5, s = bit & 31;
    uint64_t wm = P::mod(k), wa = a.v[k];
    if (k + 1 < N) { wm |= (uint64_t)P::mod(k + 1 < N ? k + 1 : 0) << 32; wa |= (uint64_t)a.v[k + 1 < N ? k + 1 : 0] << 32; }
    m.v[i] = (int32_t)((uint32_t)(wm >> s) & M30);
    g.v[i] = (int32_t)((uint32_t)(wa >> s) & M30);
    f.v[i] = m.v[i];
    d.v[i] = 0;
    e.v[i] = i == 0 ? 1 : 0;
  }
  const uint32_t m_inv30 = (0u - P::INV32) & M30;   // INV32 = -p^-1 mod 2^32
  int32_t zeta = -1;
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
    int32_t t[4];
    zeta = safegcd_divsteps_30(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
    safegcd_update_de<L>(d, e, t, m, m_inv30);
    safegcd_update_fg<L>(f, g, t);
  }
  safegcd_normalize<L>(d, f.v[L - 1], m);   // f = +-1: d = +-(aR)^-1
  Fp<P> y;
#pragma unroll
  for (int k = 0; k < N; k++) {
    const int bit = 32 * k, i = bit / 30, s = bit % 30;
    uint64_t w = (uint64_t)(uint32_t)d.v[i] >> s;
    if (i + 1 < L) w |= (uint64_t)(uint32_t)d.v[i + 1 < L ? i + 1 : 0] << (30 - s);
    if (i + 2 < L) w |= (uint64_t)(uint32_t)d.v[i + 2 < L ? i + 2 : 0] << (60 - s);
    y.v[k] = (uint32_t)w;
  }
  // y = (xR)^-1 = x^-1 R^-1;  times R^3 under the Montgomery product (which divides by R) gives x^-1 R
  const Fp<P> r3 = Fp<P>::mul(Fp<P>::r2(), Fp<P>::r2());
  return Fp<P>::mul(y, r3);
}

}  // namespace g16
