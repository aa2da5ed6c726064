// Host check of groth16_b200/csrc/fp.cuh: reads "<field> a b mul add sub inv" hex vectors (from python big ints)
// and verifies the Fp<P> implementation.  Built twice by tests/test_fp_host.py: plain host path (u64 CIOS) and
// -DG16_EMULATE_PTX (the device carry-chain algorithm with emulated PTX primitives).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <iostream>
#include "../../groth16_b200/csrc/fp.cuh"
using namespace g16;

template <class F>
static F parse(const std::string& h) {
  F r = F::zero();
  int n = (int)h.size();
  for (int i = 0; i < n; i++) {
    char c = h[n - 1 - i];
    uint32_t d = (c >= '0' && c <= '9') ? c - '0' : (c - 'a' + 10);
    r.v[i / 8] |= d << (4 * (i % 8));
  }
  return r;
}
template <class F>
static int check(const std::string& a, const std::string& b, const std::string& m, const std::string& s,
                 const std::string& d, const std::string& iv) {
  F A = parse<F>(a), B = parse<F>(b);
  int bad = 0;
  if (!(F::mul(A, B) == parse<F>(m))) { bad++; fprintf(stderr, "mul mismatch\n"); }
  if (!(F::add(A, B) == parse<F>(s))) { bad++; fprintf(stderr, "add mismatch\n"); }
  if (!(F::sub(A, B) == parse<F>(d))) { bad++; fprintf(stderr, "sub mismatch\n"); }
  if (!(F::inv(A) == parse<F>(iv))) { bad++; fprintf(stderr, "inv mismatch\n"); }
  if (!(F::add(A, F::neg(A)).is_zero())) { bad++; fprintf(stderr, "neg mismatch\n"); }
  if (!(F::to_mont(F::from_mont(A)) == A)) { bad++; fprintf(stderr, "mont roundtrip mismatch\n"); }
  return bad;
}
int main() {
  std::string f, a, b, m, s, d, iv;
  int bad = 0, n = 0;
  while (std::cin >> f >> a >> b >> m >> s >> d >> iv) {
    n++;
    if (f == "bls381_fr") bad += check<Fp<BLS381_FrP>>(a, b, m, s, d, iv);
    else if (f == "bls381_fq") bad += check<Fp<BLS381_FqP>>(a, b, m, s, d, iv);
    else if (f == "bn254_fr") bad += check<Fp<BN254_FrP>>(a, b, m, s, d, iv);
    else if (f == "bn254_fq") bad += check<Fp<BN254_FqP>>(a, b, m, s, d, iv);
    else if (f == "bls377_fr") bad += check<Fp<BLS377_FrP>>(a, b, m, s, d, iv);
    else if (f == "bls377_fq") bad += check<Fp<BLS377_FqP>>(a, b, m, s, d, iv);
    else { fprintf(stderr, "unknown field %s\n", f.c_str()); return 2; }
  }
  printf("%d vectors, %d mismatches\n", n, bad);
  return bad ? 1 : 0;
}
