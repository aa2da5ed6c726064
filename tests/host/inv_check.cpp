// Host check of groth16_b200/csrc/fp_inv.cuh: the safegcd inversion against the Fermat inversion Fp::inv for all six
// fields -- random elements, small values, powers of two, p - 1, (p +- 1) / 2, zero -- and x * inv(x) == 1.
#include <cstdio>
#include "../../groth16_b200/csrc/fp_inv.cuh"
using namespace g16;

static uint64_t seed = 99;
static uint32_t rnd() { seed = seed * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(seed >> 32); }

template <class P>
static int run(const char* name, int nrand) {
  using F = Fp<P>;
  int bad = 0, n = 0;
  auto chk = [&](const F& x) {
    n++;
    const F want = F::inv(x), got = fp_inv_safegcd<P>(x);
    if (!(got == want)) { bad++; if (bad < 5) fprintf(stderr, "%s: mismatch at case %d\n", name, n); }
    if (!x.is_zero() && !(F::mul(x, got) == F::one())) { bad++; if (bad < 5) fprintf(stderr, "%s: x * inv != 1 at case %d\n", name, n); }
  };
  chk(F::zero());
  chk(F::one());
  chk(F::neg(F::one()));
  F two = F::dbl(F::one());
  chk(two);
  chk(F::inv(two));                 // (p + 1) / 2 in the Montgomery domain
  chk(F::neg(F::inv(two)));         // (p - 1) / 2
  // raw limb patterns below the modulus: 1, 2, 2^k, 2^k - 1 as plain integers (i.e. arbitrary field elements)
  for (int k = 0; k < P::BITS - 1; k++) {
    F x = F::zero();
    x.v[k / 32] = 1u << (k % 32);
    chk(x);
    F y = F::zero();
    for (int j = 0; j <= k; j++) y.v[j / 32] |= 1u << (j % 32);
    chk(y);
  }
  {
    F x = F::modulus();             // p - 1 as a plain integer
    x.v[0] -= 1;
    chk(x);
    x.v[0] -= 1;
    chk(x);
  }
  for (int i = 0; i < nrand; i++) {
    F x;
    for (int j = 0; j < F::N; j++) x.v[j] = rnd();
    x.v[F::N - 1] &= (P::BITS % 32) ? ((1u << (P::BITS % 32 - 1)) - 1) : 0x7fffffffu;   // below 2^(BITS-1) < p
    chk(x);
    chk(F::sqr(x));
  }
  printf("%s: %d cases, %d mismatches\n", name, n, bad);
  return bad;
}

int main() {
  int bad = 0;
  bad += run<BLS381_FqP>("bls381_fq", 300);
  bad += run<BLS381_FrP>("bls381_fr", 300);
  bad += run<BN254_FqP>("bn254_fq", 300);
  bad += run<BN254_FrP>("bn254_fr", 300);
  bad += run<BLS377_FqP>("bls377_fq", 300);
  bad += run<BLS377_FrP>("bls377_fr", 300);
  printf("total %d mismatches\n", bad);
  return bad ? 1 : 0;
}
