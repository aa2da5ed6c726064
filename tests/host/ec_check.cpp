// Host check of groth16_b200/csrc/ec.cuh (plain host back-end; the same templates run on the device): XYZZ mixed / full
// addition with every exceptional case, doubling, scalar multiplication and normalisation, G1 and G2 of the three curves,
// against vectors from the big-int oracle.  Line: <curve> <g1|g2> P Q k  P+Q 2P kP   (affine coordinates, hex; "inf").
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>
#include "../../groth16_b200/csrc/ec.cuh"
using namespace g16;

template <class F>
static F parse_fp(const std::string& h) {
  F r = F::zero();
  int n = (int)h.size();
  for (int i = 0; i < n; i++) {
    char c = h[n - 1 - i];
    uint32_t d = (c >= '0' && c <= '9') ? c - '0' : (c - 'a' + 10);
    r.v[i / 8] |= d << (4 * (i % 8));
  }
  return F::to_mont(r);
}
template <class P> struct Rd1 {
  using F = Fp<P>;
  static bool point(std::istream& in, Affine<F>& p) {
    std::string a;
    in >> a;
    if (a == "inf") { p = Affine<F>::inf(); return true; }
    std::string b;
    in >> b;
    p = {parse_fp<F>(a), parse_fp<F>(b)};
    return true;
  }
};
template <class P, int NR> struct Rd2 {
  using F = Fp2<P, NR>;
  static bool point(std::istream& in, Affine<F>& p) {
    std::string a;
    in >> a;
    if (a == "inf") { p = Affine<F>::inf(); return true; }
    std::string b, c, d;
    in >> b >> c >> d;
    p = {F{parse_fp<Fp<P>>(a), parse_fp<Fp<P>>(b)}, F{parse_fp<Fp<P>>(c), parse_fp<Fp<P>>(d)}};
    return true;
  }
};
template <class F>
static bool same(const Affine<F>& a, const Affine<F>& b) { return a.x == b.x && a.y == b.y; }

template <class F, class RD>
static int run(std::istream& in) {
  Affine<F> Pp, Q, S, D, K;
  std::string ks;
  RD::point(in, Pp); RD::point(in, Q);
  in >> ks;
  RD::point(in, S); RD::point(in, D); RD::point(in, K);
  uint32_t k[8] = {0};
  for (int i = 0; i < (int)ks.size(); i++) {
    char c = ks[ks.size() - 1 - i];
    uint32_t d = (c >= '0' && c <= '9') ? c - '0' : (c - 'a' + 10);
    k[i / 8] |= d << (4 * (i % 8));
  }
  int bad = 0;
  auto chk = [&](bool ok, const char* what) { if (!ok) { bad++; fprintf(stderr, "%s mismatch\n", what); } };
  XYZZ<F> a = XYZZ<F>::from_affine(Pp);
  a.madd(Q);
  chk(same(a.to_affine(), S), "madd");
  XYZZ<F> b = XYZZ<F>::from_affine(Pp);
  b.add(XYZZ<F>::from_affine(Q));
  chk(same(b.to_affine(), S), "add");
  // non-trivial ZZ on both sides: (P + Q) + (P + Q) - P - Q ... use 2P + Q = P + (P + Q)
  XYZZ<F> c = XYZZ<F>::from_affine(Pp);
  c.add(a);
  XYZZ<F> d2 = XYZZ<F>::from_affine(Pp);
  d2.dbl_inplace();
  chk(same(d2.to_affine(), D), "dbl_inplace");
  chk(same(XYZZ<F>::dbl_affine(Pp).to_affine(), D), "dbl_affine");
  XYZZ<F> e = d2;
  e.madd(Q);
  chk(same(e.to_affine(), c.to_affine()), "2P+Q two ways");
  XYZZ<F> f = XYZZ<F>::from_affine(Pp);
  f.madd(Pp);                                    // P + P through the addition path
  chk(same(f.to_affine(), D), "madd doubling case");
  XYZZ<F> g = d2;
  g.add(d2);                                     // full add with equal operands -> doubling
  XYZZ<F> g2 = d2;
  g2.dbl_inplace();
  chk(same(g.to_affine(), g2.to_affine()), "add doubling case");
  XYZZ<F> h = XYZZ<F>::from_affine(Pp);
  h.madd(Pp, true);                              // P + (-P)
  chk(h.is_inf() && h.to_affine().is_inf(), "madd inverse case");
  XYZZ<F> h2 = d2, h3 = d2;
  h3.negate();
  h2.add(h3);
  chk(h2.is_inf(), "add inverse case");
  XYZZ<F> z = XYZZ<F>::inf();
  z.madd(Q);
  chk(same(z.to_affine(), Q), "inf + Q");
  XYZZ<F> z2 = XYZZ<F>::from_affine(Pp);
  z2.madd(Affine<F>::inf());
  z2.add(XYZZ<F>::inf());
  chk(same(z2.to_affine(), Pp), "P + inf");
  chk(same(XYZZ<F>::from_affine(Pp).mul_u32(k, 8).to_affine(), K), "mul_u32");
  return bad;
}

int main() {
  std::string curve, grp;
  int bad = 0, n = 0;
  while (std::cin >> curve >> grp) {
    n++;
    if (curve == "bls12_381") bad += grp == "g1" ? run<Fp<BLS381_FqP>, Rd1<BLS381_FqP>>(std::cin) : run<Fp2<BLS381_FqP, 1>, Rd2<BLS381_FqP, 1>>(std::cin);
    else if (curve == "bn254") bad += grp == "g1" ? run<Fp<BN254_FqP>, Rd1<BN254_FqP>>(std::cin) : run<Fp2<BN254_FqP, 1>, Rd2<BN254_FqP, 1>>(std::cin);
    else if (curve == "bls12_377") bad += grp == "g1" ? run<Fp<BLS377_FqP>, Rd1<BLS377_FqP>>(std::cin) : run<Fp2<BLS377_FqP, 5>, Rd2<BLS377_FqP, 5>>(std::cin);
    else { fprintf(stderr, "unknown curve\n"); return 2; }
  }
  printf("%d vectors, %d mismatches\n", n, bad);
  return bad ? 1 : 0;
}
