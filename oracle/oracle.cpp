// oracle/oracle.cpp -- TEST INFRASTRUCTURE ONLY.  Nothing under groth16_b200/ links, loads or calls this file.
//
// Multi-threaded CPU restatement of the Groth16 proving hot path that ark-groth16 0.5.0 (/root/reference) drives
// through its un-vendored dependencies ark-ff / ark-ec / ark-poly 0.5.0 (source not on this machine, SURVEY.md
// section 2a).  It is (i) the full-size checker for the CUDA path (bit-exact proofs at 2^20) and (ii) the timed CPU
// baseline of bench.py (`cpu_baseline`, `--impl reference`), labelled "restated ark CPU path" -- never ark itself.
//
// PARITY PINNING: the reference has no golden vectors and cannot be built here (no Rust), so
// this file is pinned against oracle/pyref.py (pure big-int, pairing-checked) by tests/test_oracle.py.
//
// Written independently of the CUDA code: 64-bit limbs with unsigned __int128 (ark-ff's representation), Jacobian
// coordinates with mixed additions (ark-ec short-Weierstrass `Projective`), Pippenger with signed digits and the
// window rule c = ceil(log2 n)*69/100 + 2 (ark-ec `ln_without_floats`), in-order radix-2 FFT (ark-poly
// Radix2EvaluationDomain semantics).  What each function follows in the reference:
//   orc_witness_map            r1cs_to_qap.rs:172-235 (evaluate_constraint :28-67, iFFT/coset FFT/pointwise/coset iFFT)
//   orc_msm_g1 / orc_msm_g2    VariableBaseMSM::msm_bigint as called at prover.rs:66,74,262
//   orc_prove                  prover.rs:26-51 + :54-132 (+ calculate_coeff :252-270)
//   orc_setup_scalars          generator.rs:47-127 + r1cs_to_qap.rs:128-170,237-247 (exponents of every key element);
//                              with orc_batch_mul_g1/g2 (generator.rs:129-183) this is a complete CPU trusted setup
// Parallelism: every MSM is split into one chunk of (scalar, base) pairs per thread, each chunk a complete serial
// Pippenger, partial results summed -- this scales to all host cores (ark-ec parallelises at least across windows;
// chunking is the stronger CPU variant, used here so that the baseline is not handicapped on many-core hosts).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

typedef uint64_t u64;
typedef unsigned __int128 u128;

// ------------------------------------------------------------------------------------------------
// parallel_for
// ------------------------------------------------------------------------------------------------
static void parallel_for(size_t ntasks, int threads, const std::function<void(size_t)>& fn) {
  if (threads <= 1 || ntasks <= 1) {
    for (size_t i = 0; i < ntasks; i++) fn(i);
    return;
  }
  std::atomic<size_t> next(0);
  int nt = (int)std::min<size_t>(threads, ntasks);
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++)
    th.emplace_back([&]() {
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= ntasks) break;
        fn(i);
      }
    });
  for (auto& x : th) x.join();
}

// ------------------------------------------------------------------------------------------------
// Fp<N, TAG>: Montgomery field with run-time modulus (one instantiation per field)
// ------------------------------------------------------------------------------------------------
template <int N, int TAG>
struct Fp {
  u64 v[N];
  static u64 MOD[N], R1[N], R2[N], INV;
  static int BITS;

  static void init(const char* hex) {
    // parse big-endian hex modulus
    for (int i = 0; i < N; i++) MOD[i] = 0;
    int len = (int)strlen(hex);
    for (int i = 0; i < len; i++) {
      char c = hex[len - 1 - i];
      u64 d = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' ? c - 'a' + 10 : c - 'A' + 10);
      MOD[i / 16] |= d << (4 * (i % 16));
    }
    BITS = 0;
    for (int i = N * 64 - 1; i >= 0; i--)
      if ((MOD[i / 64] >> (i % 64)) & 1) { BITS = i + 1; break; }
    // INV = -p^-1 mod 2^64 (Newton)
    u64 x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - MOD[0] * x;
    INV = (u64)0 - x;
    // R1 = 2^(64N) mod p by 64N modular doublings of 1; R2 = 2^(128N) mod p
    Fp t;
    for (int i = 0; i < N; i++) t.v[i] = 0;
    t.v[0] = 1;
    for (int i = 0; i < 64 * N; i++) t = raw_dbl(t);
    memcpy(R1, t.v, sizeof(R1));
    for (int i = 0; i < 64 * N; i++) t = raw_dbl(t);
    memcpy(R2, t.v, sizeof(R2));
  }
  static bool geq_mod(const u64* a) {
    for (int i = N - 1; i >= 0; i--) {
      if (a[i] > MOD[i]) return true;
      if (a[i] < MOD[i]) return false;
    }
    return true;
  }
  static void sub_mod_inplace(u64* a) {
    u64 bw = 0;
    for (int i = 0; i < N; i++) {
      u128 d = (u128)a[i] - MOD[i] - bw;
      a[i] = (u64)d;
      bw = (u64)(d >> 64) & 1;
    }
  }
  static Fp raw_dbl(const Fp& a) {  // 2a mod p for a < p (p has a spare top bit)
    Fp r;
    u64 c = 0;
    for (int i = 0; i < N; i++) {
      r.v[i] = (a.v[i] << 1) | c;
      c = a.v[i] >> 63;
    }
    if (geq_mod(r.v)) sub_mod_inplace(r.v);
    return r;
  }
  static Fp zero() { Fp r; memset(r.v, 0, sizeof(r.v)); return r; }
  static Fp one() { Fp r; memcpy(r.v, R1, sizeof(R1)); return r; }
  bool is_zero() const { u64 a = 0; for (int i = 0; i < N; i++) a |= v[i]; return a == 0; }
  bool operator==(const Fp& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
  static Fp add(const Fp& a, const Fp& b) {
    Fp r;
    u128 c = 0;
    for (int i = 0; i < N; i++) { c += (u128)a.v[i] + b.v[i]; r.v[i] = (u64)c; c >>= 64; }
    if (geq_mod(r.v)) sub_mod_inplace(r.v);
    return r;
  }
  static Fp sub(const Fp& a, const Fp& b) {
    Fp r;
    u64 bw = 0;
    for (int i = 0; i < N; i++) {
      u128 d = (u128)a.v[i] - b.v[i] - bw;
      r.v[i] = (u64)d;
      bw = (u64)(d >> 64) & 1;
    }
    if (bw) {
      u128 c = 0;
      for (int i = 0; i < N; i++) { c += (u128)r.v[i] + MOD[i]; r.v[i] = (u64)c; c >>= 64; }
    }
    return r;
  }
  static Fp neg(const Fp& a) { return a.is_zero() ? a : sub(zero(), a); }
  static Fp dbl(const Fp& a) { return add(a, a); }
  static inline __attribute__((always_inline)) Fp mul(const Fp& a, const Fp& b) {  // CIOS, no-carry variant (spare top bit)
    u64 P[N];
#pragma GCC unroll 8
    for (int i = 0; i < N; i++) P[i] = MOD[i];
    const u64 inv = INV;
    u64 t[N];
#pragma GCC unroll 8
    for (int i = 0; i < N; i++) t[i] = 0;
#pragma GCC unroll 8
    for (int i = 0; i < N; i++) {
      // t += a * b[i]
      u128 c = (u128)a.v[0] * b.v[i] + t[0];
      u64 lo0 = (u64)c;
      const u64 m = lo0 * inv;
      u128 d = (u128)m * P[0] + lo0;   // low limb becomes zero
      u64 c1 = (u64)(c >> 64), c2 = (u64)(d >> 64);
#pragma GCC unroll 8
      for (int j = 1; j < N; j++) {
        c = (u128)a.v[j] * b.v[i] + t[j] + c1;
        c1 = (u64)(c >> 64);
        d = (u128)m * P[j] + (u64)c + c2;
        c2 = (u64)(d >> 64);
        t[j - 1] = (u64)d;
      }
      t[N - 1] = c1 + c2;   // cannot overflow: modulus has a spare top bit (ark-ff's "no-carry" optimisation)
    }
    Fp r;
#pragma GCC unroll 8
    for (int i = 0; i < N; i++) r.v[i] = t[i];
    if (geq_mod(r.v)) sub_mod_inplace(r.v);
    return r;
  }
  static Fp sqr(const Fp& a) { return mul(a, a); }
  static Fp from_mont(const Fp& a) { Fp o = zero(); o.v[0] = 1; return mul(a, o); }
  static Fp to_mont(const Fp& a) { Fp r2; memcpy(r2.v, R2, sizeof(R2)); return mul(a, r2); }
  static Fp from_u64(u64 x) { Fp r = zero(); r.v[0] = x; return to_mont(r); }
  static Fp pow(const Fp& a, const u64* e, int nl) {
    Fp r = one();
    for (int i = nl * 64 - 1; i >= 0; i--) {
      r = sqr(r);
      if ((e[i / 64] >> (i % 64)) & 1) r = mul(r, a);
    }
    return r;
  }
  static Fp inv(const Fp& a) {
    u64 e[N];
    memcpy(e, MOD, sizeof(e));
    e[0] -= 2;  // p is odd and > 2: no borrow
    return pow(a, e, N);
  }
};
template <int N, int TAG> u64 Fp<N, TAG>::MOD[N];
template <int N, int TAG> u64 Fp<N, TAG>::R1[N];
template <int N, int TAG> u64 Fp<N, TAG>::R2[N];
template <int N, int TAG> u64 Fp<N, TAG>::INV;
template <int N, int TAG> int Fp<N, TAG>::BITS;

template <class B, int NR>
struct Fp2 {
  B c0, c1;
  static Fp2 zero() { return {B::zero(), B::zero()}; }
  static Fp2 one() { return {B::one(), B::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
  static Fp2 add(const Fp2& a, const Fp2& b) { return {B::add(a.c0, b.c0), B::add(a.c1, b.c1)}; }
  static Fp2 sub(const Fp2& a, const Fp2& b) { return {B::sub(a.c0, b.c0), B::sub(a.c1, b.c1)}; }
  static Fp2 neg(const Fp2& a) { return {B::neg(a.c0), B::neg(a.c1)}; }
  static Fp2 dbl(const Fp2& a) { return {B::dbl(a.c0), B::dbl(a.c1)}; }
  static B mulnr(const B& a) {
    if (NR == 1) return a;
    B t = B::dbl(B::dbl(a));  // 4a
    return B::add(t, a);      // 5a   (NR is 1 or 5)
  }
  static Fp2 mul(const Fp2& a, const Fp2& b) {
    B v0 = B::mul(a.c0, b.c0), v1 = B::mul(a.c1, b.c1);
    B s = B::mul(B::add(a.c0, a.c1), B::add(b.c0, b.c1));
    return {B::sub(v0, mulnr(v1)), B::sub(B::sub(s, v0), v1)};
  }
  static Fp2 sqr(const Fp2& a) { return mul(a, a); }
  static Fp2 inv(const Fp2& a) {
    B n = B::add(B::sqr(a.c0), mulnr(B::sqr(a.c1)));
    B ni = B::inv(n);
    return {B::mul(a.c0, ni), B::neg(B::mul(a.c1, ni))};
  }
};

// ------------------------------------------------------------------------------------------------
// Jacobian points, a = 0
// ------------------------------------------------------------------------------------------------
template <class F>
struct Aff {
  F x, y;
  bool is_inf() const { return x.is_zero() && y.is_zero(); }  // ABI encoding of the identity
};
template <class F>
struct Jac {
  F X, Y, Z;
  static Jac inf() { return {F::one(), F::one(), F::zero()}; }
  bool is_inf() const { return Z.is_zero(); }
  void dbl() {  // dbl-2009-l
    if (is_inf()) return;
    F A = F::sqr(X), B = F::sqr(Y), C = F::sqr(B);
    F t = F::add(X, B);
    F D = F::dbl(F::sub(F::sub(F::sqr(t), A), C));
    F E = F::add(F::dbl(A), A);
    F Fv = F::sqr(E);
    F Z3 = F::dbl(F::mul(Y, Z));
    X = F::sub(Fv, F::dbl(D));
    F C8 = F::dbl(F::dbl(F::dbl(C)));
    Y = F::sub(F::mul(E, F::sub(D, X)), C8);
    Z = Z3;
  }
  void madd(const Aff<F>& p, bool negate) {  // madd-2007-bl
    if (p.is_inf()) return;
    F py = negate ? F::neg(p.y) : p.y;
    if (is_inf()) { X = p.x; Y = py; Z = F::one(); return; }
    F Z1Z1 = F::sqr(Z);
    F U2 = F::mul(p.x, Z1Z1);
    F S2 = F::mul(F::mul(py, Z), Z1Z1);
    if (U2 == X) {
      if (S2 == Y) { dbl(); return; }
      *this = inf();
      return;
    }
    F H = F::sub(U2, X);
    F HH = F::sqr(H);
    F I = F::dbl(F::dbl(HH));
    F J = F::mul(H, I);
    F r = F::dbl(F::sub(S2, Y));
    F V = F::mul(X, I);
    F X3 = F::sub(F::sub(F::sqr(r), J), F::dbl(V));
    F Y3 = F::sub(F::mul(r, F::sub(V, X3)), F::dbl(F::mul(Y, J)));
    F Z3 = F::sub(F::sub(F::sqr(F::add(Z, H)), Z1Z1), HH);
    X = X3; Y = Y3; Z = Z3;
  }
  void add(const Jac& q) {  // add-2007-bl
    if (q.is_inf()) return;
    if (is_inf()) { *this = q; return; }
    F Z1Z1 = F::sqr(Z), Z2Z2 = F::sqr(q.Z);
    F U1 = F::mul(X, Z2Z2), U2 = F::mul(q.X, Z1Z1);
    F S1 = F::mul(F::mul(Y, q.Z), Z2Z2), S2 = F::mul(F::mul(q.Y, Z), Z1Z1);
    if (U1 == U2) {
      if (S1 == S2) { dbl(); return; }
      *this = inf();
      return;
    }
    F H = F::sub(U2, U1);
    F I = F::sqr(F::dbl(H));
    F J = F::mul(H, I);
    F r = F::dbl(F::sub(S2, S1));
    F V = F::mul(U1, I);
    F X3 = F::sub(F::sub(F::sqr(r), J), F::dbl(V));
    F Y3 = F::sub(F::mul(r, F::sub(V, X3)), F::dbl(F::mul(S1, J)));
    F Z3 = F::mul(F::sub(F::sub(F::sqr(F::add(Z, q.Z)), Z1Z1), Z2Z2), H);
    X = X3; Y = Y3; Z = Z3;
  }
  void neg() { Y = F::neg(Y); }
  Jac mul_bits(const u64* k, int nl) const {
    Jac r = inf();
    for (int i = nl * 64 - 1; i >= 0; i--) {
      r.dbl();
      if ((k[i / 64] >> (i % 64)) & 1) r.add(*this);
    }
    return r;
  }
  Aff<F> to_affine() const {
    if (is_inf()) return {F::zero(), F::zero()};
    F zi = F::inv(Z), zi2 = F::sqr(zi);
    return {F::mul(X, zi2), F::mul(Y, F::mul(zi2, zi))};
  }
  static Jac from_affine(const Aff<F>& p) { return p.is_inf() ? inf() : Jac{p.x, p.y, F::one()}; }
};

// ------------------------------------------------------------------------------------------------
// Pippenger (serial, one chunk) -- signed digits, window rule of ark-ec
// ------------------------------------------------------------------------------------------------
static int ark_window(size_t n) {
  if (n < 32) return 3;
  int lg = 0;
  while (((size_t)1 << lg) < n) lg++;
  return lg * 69 / 100 + 2;
}
template <class F>
static Jac<F> pippenger_serial(const Aff<F>* bases, const u64* scalars /* 4 limbs each, canonical */, size_t n, int bits) {
  if (n == 0) return Jac<F>::inf();
  const int c = ark_window(n);
  const int W = (bits + 1 + c - 1) / c;
  const size_t B = (size_t)1 << (c - 1);
  // signed digits
  std::vector<int32_t> digits(n * W);
  for (size_t i = 0; i < n; i++) {
    const u64* s = scalars + 4 * i;
    u64 carry = 0;
    for (int w = 0; w < W; w++) {
      int bit = w * c, li = bit / 64, sh = bit % 64;
      u64 raw = 0;
      if (li < 4) {
        raw = s[li] >> sh;
        if (sh + c > 64 && li + 1 < 4) raw |= s[li + 1] << (64 - sh);
        raw &= ((u64)1 << c) - 1;
      }
      raw += carry;
      carry = 0;
      int32_t d = (int32_t)raw;
      if (raw > B) { d = (int32_t)raw - (int32_t)((u64)1 << c); carry = 1; }
      digits[i * W + w] = d;
    }
  }
  Jac<F> total = Jac<F>::inf();
  std::vector<Jac<F>> buckets(B);
  for (int w = W - 1; w >= 0; w--) {
    for (auto& b : buckets) b = Jac<F>::inf();
    for (size_t i = 0; i < n; i++) {
      int32_t d = digits[i * W + w];
      if (d > 0) buckets[d - 1].madd(bases[i], false);
      else if (d < 0) buckets[-d - 1].madd(bases[i], true);
    }
    Jac<F> running = Jac<F>::inf(), sum = Jac<F>::inf();
    for (size_t b = B; b-- > 0;) {
      running.add(buckets[b]);
      sum.add(running);
    }
    for (int k = 0; k < c; k++) total.dbl();
    total.add(sum);
  }
  return total;
}
// ark-ec 0.5.0 `msm_bigint` parallelises ACROSS WINDOWS (`ark_std::cfg_into_iter!(window_starts)`: every window walks
// all n pairs serially, window sums are combined at the end), so its speed-up saturates at W = ceil(256 / c) threads.
// Kept selectable (orc_set_msm_mode(1)) so that bench.py can print both CPU variants side by side; the default stays
// the chunk-parallel one below, which is the stronger baseline on many-core hosts.
static std::atomic<int> g_msm_mode(0);   // 0 = chunk-parallel (default), 1 = window-parallel (ark-ec's scheme)
template <class F>
static Jac<F> msm_window_parallel(const Aff<F>* bases, const u64* scalars, size_t n, int bits, int threads) {
  if (n == 0) return Jac<F>::inf();
  const int c = ark_window(n);
  const int W = (bits + 1 + c - 1) / c;
  const size_t B = (size_t)1 << (c - 1);
  std::vector<int32_t> digits(n * W);
  const size_t grain = 1 << 14, tasks = (n + grain - 1) / grain;
  parallel_for(tasks, threads, [&](size_t t) {
    for (size_t i = t * grain; i < std::min(n, (t + 1) * grain); i++) {
      const u64* s = scalars + 4 * i;
      u64 carry = 0;
      for (int w = 0; w < W; w++) {
        int bit = w * c, li = bit / 64, sh = bit % 64;
        u64 raw = 0;
        if (li < 4) {
          raw = s[li] >> sh;
          if (sh + c > 64 && li + 1 < 4) raw |= s[li + 1] << (64 - sh);
          raw &= ((u64)1 << c) - 1;
        }
        raw += carry;
        carry = 0;
        int32_t d = (int32_t)raw;
        if (raw > B) { d = (int32_t)raw - (int32_t)((u64)1 << c); carry = 1; }
        digits[i * W + w] = d;
      }
    }
  });
  std::vector<Jac<F>> wsum(W);
  parallel_for((size_t)W, threads, [&](size_t w) {
    std::vector<Jac<F>> buckets(B, Jac<F>::inf());
    for (size_t i = 0; i < n; i++) {
      int32_t d = digits[i * W + w];
      if (d > 0) buckets[d - 1].madd(bases[i], false);
      else if (d < 0) buckets[-d - 1].madd(bases[i], true);
    }
    Jac<F> running = Jac<F>::inf(), sum = Jac<F>::inf();
    for (size_t b = B; b-- > 0;) { running.add(buckets[b]); sum.add(running); }
    wsum[w] = sum;
  });
  Jac<F> total = Jac<F>::inf();
  for (int w = W - 1; w >= 0; w--) {
    for (int k = 0; k < c; k++) total.dbl();
    total.add(wsum[w]);
  }
  return total;
}
template <class F>
static Jac<F> msm_parallel(const Aff<F>* bases, const u64* scalars, size_t n, int bits, int threads) {
  if (n == 0) return Jac<F>::inf();
  if (g_msm_mode.load() == 1) return msm_window_parallel<F>(bases, scalars, n, bits, threads);
  size_t chunks = std::max<size_t>(1, std::min<size_t>((size_t)threads, n / 1024 + 1));
  std::vector<Jac<F>> part(chunks);
  parallel_for(chunks, threads, [&](size_t t) {
    size_t lo = n * t / chunks, hi = n * (t + 1) / chunks;
    part[t] = pippenger_serial<F>(bases + lo, scalars + 4 * lo, hi - lo, bits);
  });
  Jac<F> r = Jac<F>::inf();
  for (auto& p : part) r.add(p);
  return r;
}


// Fixed-base batch multiplication out[i] = scalars[i] * g (BatchMulPreprocessing::batch_mul, generator.rs:129-183):
// 8-bit windows, table of 32 x 255 Jacobian multiples, one inversion per output.  scalars: Montgomery Fr.
template <class F, class Fr>
static void batch_mul(const Aff<F>& g, const Fr* scalars, size_t n, Aff<F>* out, int threads) {
  std::vector<Jac<F>> table(32 * 255);
  Jac<F> base = Jac<F>::from_affine(g);
  for (int w = 0; w < 32; w++) {
    Jac<F> acc = base;
    for (int d = 1; d <= 255; d++) { table[w * 255 + d - 1] = acc; acc.add(base); }
    base = acc;  // 256 * base
  }
  const size_t grain = 256, tasks = (n + grain - 1) / grain;
  parallel_for(tasks, threads, [&](size_t t) {
    for (size_t i = t * grain; i < std::min(n, (t + 1) * grain); i++) {
      Fr c = Fr::from_mont(scalars[i]);
      Jac<F> acc = Jac<F>::inf();
      for (int w = 0; w < 32; w++) {
        unsigned d = (unsigned)(c.v[w / 8] >> (8 * (w % 8))) & 0xff;
        if (d) acc.add(table[w * 255 + d - 1]);
      }
      out[i] = acc.to_affine();
    }
  });
}

// ------------------------------------------------------------------------------------------------
// Radix-2 domain (ark-poly semantics: natural order in/out)
// ------------------------------------------------------------------------------------------------
template <class Fr>
struct Domain {
  int L;
  size_t n;
  Fr omega, omega_inv, n_inv, g, g_inv;
  std::vector<Fr> tw, tw_inv;  // omega^i, i < n/2
};
// a[i] *= c0 * base^i, chunked over threads
template <class Fr>
static void scale_powers(std::vector<Fr>& a, const Fr& c0, const Fr& base, bool mul_base, int threads) {
  const size_t n = a.size(), grain = 1 << 13, tasks = (n + grain - 1) / grain;
  parallel_for(tasks, tasks > 1 ? threads : 1, [&](size_t t) {
    size_t lo = t * grain, hi = std::min(n, lo + grain);
    Fr p = c0;
    if (mul_base) { u64 e[1] = {(u64)lo}; p = Fr::mul(p, Fr::pow(base, e, 1)); }
    for (size_t i = lo; i < hi; i++) { a[i] = Fr::mul(a[i], p); if (mul_base) p = Fr::mul(p, base); }
  });
}
static inline size_t bitrev(size_t i, int L) {
  size_t j = 0;
  for (int b = 0; b < L; b++) j |= ((i >> b) & 1) << (L - 1 - b);
  return j;
}
template <class Fr>
static void fft_core(std::vector<Fr>& a, const std::vector<Fr>& tw, int L, int threads) {
  const size_t n = a.size();
  {
    const size_t grain = 1 << 14, tasks = (n + grain - 1) / grain;
    parallel_for(tasks, tasks > 1 ? threads : 1, [&](size_t t) {
      for (size_t i = t * grain; i < std::min(n, (t + 1) * grain); i++) { size_t j = bitrev(i, L); if (i < j) std::swap(a[i], a[j]); }
    });
  }
  for (int s = 1; s <= L; s++) {
    const size_t m = (size_t)1 << s, half = m >> 1, stride = n / m;
    const size_t nbf = n / 2;
    const size_t grain = 1 << 12;
    const size_t tasks = (nbf + grain - 1) / grain;
    parallel_for(tasks, tasks > 1 ? threads : 1, [&](size_t t) {
      size_t lo = t * grain, hi = std::min(nbf, lo + grain);
      for (size_t bf = lo; bf < hi; bf++) {
        size_t blk = bf >> (s - 1), j = bf & (half - 1);
        size_t i0 = blk * m + j, i1 = i0 + half;
        Fr v = Fr::mul(a[i1], tw[j * stride]);
        Fr u = a[i0];
        a[i0] = Fr::add(u, v);
        a[i1] = Fr::sub(u, v);
      }
    });
  }
}

// ------------------------------------------------------------------------------------------------
// Curve bundle
// ------------------------------------------------------------------------------------------------
struct Csr { const uint32_t* row_ptr; const uint32_t* col; const u64* val; };
struct PkDesc {
  const u64* a_query; u64 a_len; const u64* b_g1_query; u64 b_g1_len; const u64* b_g2_query; u64 b_g2_len;
  const u64* h_query; u64 h_len; const u64* l_query; u64 l_len;
  const u64* alpha_g1; const u64* beta_g1; const u64* delta_g1; const u64* beta_g2; const u64* delta_g2;
};

template <class Fr, class Fq, int NR, int GEN, int TWO_ADICITY>
struct Curve {
  using Fq2 = Fp2<Fq, NR>;
  using A1 = Aff<Fq>;
  using A2 = Aff<Fq2>;
  using J1 = Jac<Fq>;
  using J2 = Jac<Fq2>;

  static const Domain<Fr>& cached_domain(int L, int threads) {
    static Domain<Fr> cache[64];
    static bool have[64] = {};
    if (!have[L]) { cache[L] = make_domain(L, threads); have[L] = true; }
    return cache[L];
  }
  static Domain<Fr> make_domain(int L, int threads) {
    Domain<Fr> d;
    d.L = L;
    d.n = (size_t)1 << L;
    // two_adic_root = GEN^((r-1)/2^s); omega = root^(2^(s-L))
    u64 e[4];
    memcpy(e, Fr::MOD, sizeof(e));
    e[0] -= 1;
    // shift right by TWO_ADICITY
    for (int k = 0; k < TWO_ADICITY; k++) {
      for (int i = 0; i < 4; i++) e[i] = (e[i] >> 1) | (i + 1 < 4 ? e[i + 1] << 63 : 0);
    }
    d.g = Fr::from_u64(GEN);
    Fr root = Fr::pow(d.g, e, 4);
    for (int i = 0; i < TWO_ADICITY - L; i++) root = Fr::sqr(root);
    d.omega = root;
    d.omega_inv = Fr::inv(root);
    d.n_inv = Fr::inv(Fr::from_u64((u64)d.n));
    d.g_inv = Fr::inv(d.g);
    size_t half = d.n > 1 ? d.n / 2 : 1;
    d.tw.resize(half);
    d.tw_inv.resize(half);
    Fr w = Fr::one(), wi = Fr::one();
    for (size_t i = 0; i < half; i++) { d.tw[i] = w; d.tw_inv[i] = wi; w = Fr::mul(w, d.omega); wi = Fr::mul(wi, d.omega_inv); }
    (void)threads;
    return d;
  }
  static void fft(const Domain<Fr>& d, std::vector<Fr>& a, bool coset, int threads) {
    if (coset) scale_powers(a, Fr::one(), d.g, true, threads);
    fft_core(a, d.tw, d.L, threads);
  }
  static void ifft(const Domain<Fr>& d, std::vector<Fr>& a, bool coset, int threads) {
    fft_core(a, d.tw_inv, d.L, threads);
    scale_powers(a, d.n_inv, d.g_inv, coset, threads);
  }
  // r1cs_to_qap.rs:201-234 from the three evaluation vectors
  static void witness_map_evals(const Domain<Fr>& d, std::vector<Fr>& a, std::vector<Fr>& b, std::vector<Fr>& c, int threads) {
    ifft(d, a, false, threads); ifft(d, b, false, threads);
    fft(d, a, true, threads); fft(d, b, true, threads);
    ifft(d, c, false, threads); fft(d, c, true, threads);
    Fr gn = d.g;
    for (int i = 0; i < d.L; i++) gn = Fr::sqr(gn);
    Fr zinv = Fr::inv(Fr::sub(gn, Fr::one()));
    for (size_t i = 0; i < d.n; i++) a[i] = Fr::mul(Fr::sub(Fr::mul(a[i], b[i]), c[i]), zinv);
    ifft(d, a, true, threads);
  }
  // r1cs_to_qap.rs:172-199,213-218
  static int witness_map(uint32_t ni, uint32_t nc, uint32_t nw, const Csr* A, const Csr* B, const Csr* C, const u64* z_, u64* h_out, int threads) {
    size_t need = (size_t)nc + ni;
    int L = 0;
    while (((size_t)1 << L) < need) L++;
    if (L > TWO_ADICITY) return 1;
    const Domain<Fr>& d = cached_domain(L, threads);
    const Fr* z = reinterpret_cast<const Fr*>(z_);
    std::vector<Fr> a(d.n, Fr::zero()), b(d.n, Fr::zero()), c(d.n, Fr::zero());
    const Csr* ms[3] = {A, B, C};
    std::vector<Fr>* outs[3] = {&a, &b, &c};
    const size_t grain = 4096, tasks = (nc + grain - 1) / grain;
    parallel_for(tasks, threads, [&](size_t t) {
      size_t lo = t * grain, hi = std::min<size_t>(nc, lo + grain);
      for (int m = 0; m < 3; m++) {
        const Fr* vals = reinterpret_cast<const Fr*>(ms[m]->val);
        for (size_t i = lo; i < hi; i++) {
          Fr acc = Fr::zero();
          for (uint32_t e = ms[m]->row_ptr[i]; e < ms[m]->row_ptr[i + 1]; e++) acc = Fr::add(acc, Fr::mul(vals[e], z[ms[m]->col[e]]));
          (*outs[m])[i] = acc;
        }
      }
    });
    for (uint32_t i = 0; i < ni; i++) a[nc + i] = z[i];
    (void)nw;
    witness_map_evals(d, a, b, c, threads);
    memcpy(h_out, a.data(), d.n * sizeof(Fr));
    return 0;
  }
  // Exponents of the proving / verifying key from explicit toxic waste: generator.rs:47-127 with
  // R1CSToQAP::instance_map_with_evaluation (r1cs_to_qap.rs:128-170) and h_query_scalars (:237-247).
  //   qa[i] = a_i(tau), qb[i] = b_i(tau)                      i < ni + nw      (a_query / b_g1_query / b_g2_query exponents)
  //   gabc[i] = (beta a_i + alpha b_i + c_i)(tau) / gamma      i < ni           (generator.rs:113-117)
  //   lq[i]   = (beta a_i + alpha b_i + c_i)(tau) / delta      ni <= i          (generator.rs:119-123)
  //   hs[i]   = tau^i Z(tau) / delta                           i < n - 1        (generator.rs:168 via :237-247)
  // All Montgomery Fr.  Returns 1 for PolynomialDegreeTooLarge, 2 when gamma/delta/Z(tau) is not invertible.
  static int setup_scalars(uint32_t ni, uint32_t nc, uint32_t nw, const Csr* A, const Csr* B, const Csr* C, const u64* toxic /* alpha,beta,gamma,delta,tau */,
                           u64* qa_, u64* qb_, u64* lq_, u64* hs_, u64* gabc_, int threads) {
    size_t need = (size_t)nc + ni;
    int L = 0;
    while (((size_t)1 << L) < need) L++;
    if (L > TWO_ADICITY) return 1;
    const Domain<Fr>& d = cached_domain(L, threads);
    const Fr* tx = reinterpret_cast<const Fr*>(toxic);
    const Fr alpha = tx[0], beta = tx[1], gamma = tx[2], delta = tx[3], tau = tx[4];
    if (gamma.is_zero() || delta.is_zero()) return 2;
    Fr tn = tau;
    for (int i = 0; i < L; i++) tn = Fr::sqr(tn);
    const Fr zt = Fr::sub(tn, Fr::one());            // domain.evaluate_vanishing_polynomial(t), r1cs_to_qap.rs:143
    if (zt.is_zero()) return 2;                       // tau inside the domain: the reference's sampling never hits it
    // evaluate_all_lagrange_coefficients(t): u_i = Z(t) w^i / (n (t - w^i)), one batch inversion per chunk of the domain
    const size_t n = d.n, nv = (size_t)ni + nw;
    std::vector<Fr> u(n);
    const Fr zn = Fr::mul(zt, d.n_inv);
    {
      const size_t grain = 1 << 12, tasks = (n + grain - 1) / grain;
      parallel_for(tasks, threads, [&](size_t t) {
        const size_t lo = t * grain, hi = std::min(n, lo + grain);
        u64 e[1] = {(u64)lo};
        Fr w = Fr::pow(d.omega, e, 1);
        std::vector<Fr> den(hi - lo), pref(hi - lo), ws(hi - lo);
        Fr acc = Fr::one();
        for (size_t i = lo; i < hi; i++) { ws[i - lo] = w; den[i - lo] = Fr::sub(tau, w); pref[i - lo] = acc; acc = Fr::mul(acc, den[i - lo]); w = Fr::mul(w, d.omega); }
        Fr ai = Fr::inv(acc);
        for (size_t k = hi - lo; k-- > 0;) { Fr di = Fr::mul(ai, pref[k]); ai = Fr::mul(ai, den[k]); u[lo + k] = Fr::mul(Fr::mul(zn, ws[k]), di); }
      });
    }
    std::vector<Fr> qa(nv, Fr::zero()), qb(nv, Fr::zero()), qc(nv, Fr::zero());
    for (uint32_t i = 0; i < ni; i++) qa[i] = u[nc + i];                                 // r1cs_to_qap.rs:150-155
    const Csr* ms[3] = {A, B, C};
    std::vector<Fr>* outs[3] = {&qa, &qb, &qc};
    parallel_for(3, threads, [&](size_t m) {                                             // r1cs_to_qap.rs:157-167 (scatter: serial per matrix)
      const Fr* vals = reinterpret_cast<const Fr*>(ms[m]->val);
      for (uint32_t i = 0; i < nc; i++)
        for (uint32_t e = ms[m]->row_ptr[i]; e < ms[m]->row_ptr[i + 1]; e++) {
          Fr& dst = (*outs[m])[ms[m]->col[e]];
          dst = Fr::add(dst, Fr::mul(u[i], vals[e]));
        }
    });
    const Fr gi = Fr::inv(gamma), di = Fr::inv(delta);
    Fr* qa_o = reinterpret_cast<Fr*>(qa_); Fr* qb_o = reinterpret_cast<Fr*>(qb_); Fr* lq = reinterpret_cast<Fr*>(lq_);
    Fr* hs = reinterpret_cast<Fr*>(hs_); Fr* gabc = reinterpret_cast<Fr*>(gabc_);
    for (size_t i = 0; i < nv; i++) {
      const Fr t = Fr::add(Fr::add(Fr::mul(beta, qa[i]), Fr::mul(alpha, qb[i])), qc[i]);
      if (i < ni) gabc[i] = Fr::mul(t, gi); else lq[i - ni] = Fr::mul(t, di);
      qa_o[i] = qa[i];
      qb_o[i] = qb[i];
    }
    Fr p = Fr::mul(zt, di);
    for (size_t i = 0; i + 1 < n; i++) { hs[i] = p; p = Fr::mul(p, tau); }
    return 0;
  }
  static void to_bigints(const Fr* in, size_t n, std::vector<u64>& out, int threads) {
    out.resize(n * 4);
    const size_t grain = 8192, tasks = (n + grain - 1) / grain;
    parallel_for(tasks, threads, [&](size_t t) {
      for (size_t i = t * grain; i < std::min(n, (t + 1) * grain); i++) { Fr c = Fr::from_mont(in[i]); memcpy(&out[4 * i], c.v, 32); }
    });
  }
  // prover.rs:26-51 + 54-132
  static int prove(const PkDesc* pk, uint32_t ni, uint32_t nc, uint32_t nw, const Csr* A, const Csr* B, const Csr* C, const u64* z_,
                   const u64* r_, const u64* s_, u64* proof, int threads, double* t_ms) {
    auto T0 = std::chrono::steady_clock::now();
    size_t need = (size_t)nc + ni;
    int L = 0;
    while (((size_t)1 << L) < need) L++;
    size_t n = (size_t)1 << L;
    std::vector<u64> hbuf(n * 4);
    int rc = witness_map(ni, nc, nw, A, B, C, z_, hbuf.data(), threads);
    if (rc) return rc;
    auto T1 = std::chrono::steady_clock::now();
    const Fr* z = reinterpret_cast<const Fr*>(z_);
    const Fr r = *reinterpret_cast<const Fr*>(r_), s = *reinterpret_cast<const Fr*>(s_);
    std::vector<u64> hs, zs;
    to_bigints(reinterpret_cast<const Fr*>(hbuf.data()), n, hs, threads);
    to_bigints(z, (size_t)ni + nw, zs, threads);
    const int bits = Fr::BITS;
    auto msm1 = [&](const u64* bases, size_t blen, const u64* sc, size_t slen) { return msm_parallel<Fq>(reinterpret_cast<const A1*>(bases), sc, std::min(blen, slen), bits, threads); };
    auto msm2 = [&](const u64* bases, size_t blen, const u64* sc, size_t slen) { return msm_parallel<Fq2>(reinterpret_cast<const A2*>(bases), sc, std::min(blen, slen), bits, threads); };
    const size_t nz1 = (size_t)ni + nw - 1;
    constexpr size_t G1L = sizeof(A1) / 8, G2L = sizeof(A2) / 8;
    J1 h_acc = msm1(pk->h_query, pk->h_len, hs.data(), n);                                        // prover.rs:66
    J1 l_acc = msm1(pk->l_query, pk->l_len, zs.data() + 4 * (size_t)ni, nw);                      // prover.rs:74
    u64 rk[4], sk[4], rsk[4];
    { Fr c = Fr::from_mont(r); memcpy(rk, c.v, 32); c = Fr::from_mont(s); memcpy(sk, c.v, 32); c = Fr::from_mont(Fr::mul(r, s)); memcpy(rsk, c.v, 32); }
    const A1 delta1 = *reinterpret_cast<const A1*>(pk->delta_g1);
    J1 d1 = J1::from_affine(delta1);
    J1 rsd = d1.mul_bits(rsk, 4);                                                                 // prover.rs:76
    J1 g_a = d1.mul_bits(rk, 4);                                                                  // prover.rs:90-92
    g_a.madd(*reinterpret_cast<const A1*>(pk->a_query), false);
    g_a.add(msm1(pk->a_query + G1L, pk->a_len - 1, zs.data() + 4, nz1));
    g_a.madd(*reinterpret_cast<const A1*>(pk->alpha_g1), false);
    J1 s_g_a = g_a.mul_bits(sk, 4);
    J1 g1_b = J1::inf();
    if (!r.is_zero()) {                                                                           // prover.rs:98-108
      g1_b = d1.mul_bits(sk, 4);
      g1_b.madd(*reinterpret_cast<const A1*>(pk->b_g1_query), false);
      g1_b.add(msm1(pk->b_g1_query + G1L, pk->b_g1_len - 1, zs.data() + 4, nz1));
      g1_b.madd(*reinterpret_cast<const A1*>(pk->beta_g1), false);
    }
    J2 g2_b = J2::from_affine(*reinterpret_cast<const A2*>(pk->delta_g2)).mul_bits(sk, 4);        // prover.rs:112-113
    g2_b.madd(*reinterpret_cast<const A2*>(pk->b_g2_query), false);
    g2_b.add(msm2(pk->b_g2_query + G2L, pk->b_g2_len - 1, zs.data() + 4, nz1));
    g2_b.madd(*reinterpret_cast<const A2*>(pk->beta_g2), false);
    J1 r_g1_b = g1_b.mul_bits(rk, 4);
    J1 g_c = s_g_a;                                                                               // prover.rs:119-124
    g_c.add(r_g1_b);
    rsd.neg();
    g_c.add(rsd);
    g_c.add(l_acc);
    g_c.add(h_acc);
    A1 pa = g_a.to_affine(); A2 pb = g2_b.to_affine(); A1 pc = g_c.to_affine();
    memcpy(proof, &pa, sizeof(A1));
    memcpy(proof + G1L, &pb, sizeof(A2));
    memcpy(proof + G1L + G2L, &pc, sizeof(A1));
    auto T2 = std::chrono::steady_clock::now();
    if (t_ms) { t_ms[0] = std::chrono::duration<double, std::milli>(T1 - T0).count(); t_ms[1] = std::chrono::duration<double, std::milli>(T2 - T1).count(); }
    return 0;
  }
  template <class F>
  static void store_proj(u64* out, const Jac<F>& p) {
    Aff<F> a = p.to_affine();
    F one = F::one(), zero = F::zero();
    const size_t w = sizeof(F) / 8;
    if (p.is_inf()) { memcpy(out, &one, sizeof(F)); memcpy(out + w, &one, sizeof(F)); memcpy(out + 2 * w, &zero, sizeof(F)); }
    else { memcpy(out, &a.x, sizeof(F)); memcpy(out + w, &a.y, sizeof(F)); memcpy(out + 2 * w, &one, sizeof(F)); }
  }
};

typedef Fp<4, 0> Fr381; typedef Fp<6, 1> Fq381;
typedef Fp<4, 2> Fr254; typedef Fp<4, 3> Fq254;
typedef Fp<4, 4> Fr377; typedef Fp<6, 5> Fq377;
typedef Curve<Fr381, Fq381, 1, 7, 32> C381;
typedef Curve<Fr254, Fq254, 1, 5, 28> C254;
typedef Curve<Fr377, Fq377, 5, 22, 47> C377;

static bool g_init = false;
static void init_all() {
  if (g_init) return;
  Fr381::init("73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001");
  Fq381::init("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab");
  Fr254::init("30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001");
  Fq254::init("30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47");
  Fr377::init("12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001");
  Fq377::init("1ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001");
  g_init = true;
}

#define DISPATCH(curve, ...)         \
  switch (curve) {                   \
    case 0: { using C = C381; __VA_ARGS__; } \
    case 1: { using C = C254; __VA_ARGS__; } \
    case 2: { using C = C377; __VA_ARGS__; } \
    default: return 2;               \
  }

extern "C" {
int orc_hw_threads() { return (int)std::thread::hardware_concurrency(); }

int orc_ntt(int curve, uint32_t log_n, int inverse, int coset, u64* inout, int threads) {
  init_all();
  DISPATCH(curve, {
    using Fr = decltype(C::make_domain(0, 1).omega);
    const auto& d = C::cached_domain((int)log_n, threads);
    std::vector<Fr> a(d.n);
    memcpy(a.data(), inout, d.n * sizeof(Fr));
    if (inverse) C::ifft(d, a, coset != 0, threads); else C::fft(d, a, coset != 0, threads);
    memcpy(inout, a.data(), d.n * sizeof(Fr));
    return 0;
  })
}
int orc_witness_map(int curve, uint32_t ni, uint32_t nc, uint32_t nw, const Csr* a, const Csr* b, const Csr* c, const u64* z, u64* h, int threads) {
  init_all();
  DISPATCH(curve, { return C::witness_map(ni, nc, nw, a, b, c, z, h, threads); })
}
int orc_msm_g1(int curve, const u64* bases, const u64* scalars, u64 n, u64* out, int threads) {
  init_all();
  DISPATCH(curve, {
    using A1 = typename C::A1;
    using Fq = decltype(A1().x);
    using Fr = decltype(C::make_domain(0, 1).omega);
    auto r = msm_parallel<Fq>(reinterpret_cast<const A1*>(bases), scalars, n, Fr::BITS, threads);
    C::template store_proj<Fq>(out, r);
    return 0;
  })
}
int orc_msm_g2(int curve, const u64* bases, const u64* scalars, u64 n, u64* out, int threads) {
  init_all();
  DISPATCH(curve, {
    using A2 = typename C::A2;
    using Fq2 = decltype(A2().x);
    using Fr = decltype(C::make_domain(0, 1).omega);
    auto r = msm_parallel<Fq2>(reinterpret_cast<const A2*>(bases), scalars, n, Fr::BITS, threads);
    C::template store_proj<Fq2>(out, r);
    return 0;
  })
}
int orc_batch_mul_g1(int curve, const u64* g, const u64* scalars, u64 n, u64* out, int threads) {
  init_all();
  DISPATCH(curve, {
    using A1 = typename C::A1;
    using Fq = decltype(A1().x);
    using Fr = decltype(C::make_domain(0, 1).omega);
    batch_mul<Fq, Fr>(*reinterpret_cast<const A1*>(g), reinterpret_cast<const Fr*>(scalars), n, reinterpret_cast<A1*>(out), threads);
    return 0;
  })
}
int orc_batch_mul_g2(int curve, const u64* g, const u64* scalars, u64 n, u64* out, int threads) {
  init_all();
  DISPATCH(curve, {
    using A2 = typename C::A2;
    using Fq2 = decltype(A2().x);
    using Fr = decltype(C::make_domain(0, 1).omega);
    batch_mul<Fq2, Fr>(*reinterpret_cast<const A2*>(g), reinterpret_cast<const Fr*>(scalars), n, reinterpret_cast<A2*>(out), threads);
    return 0;
  })
}
int orc_set_msm_mode(int mode) { int old = g_msm_mode.exchange(mode == 1 ? 1 : 0); return old; }
int orc_setup_scalars(int curve, uint32_t ni, uint32_t nc, uint32_t nw, const Csr* a, const Csr* b, const Csr* c, const u64* toxic,
                      u64* qa, u64* qb, u64* lq, u64* hs, u64* gabc, int threads) {
  init_all();
  DISPATCH(curve, { return C::setup_scalars(ni, nc, nw, a, b, c, toxic, qa, qb, lq, hs, gabc, threads); })
}
int orc_prove(int curve, const PkDesc* pk, uint32_t ni, uint32_t nc, uint32_t nw, const Csr* a, const Csr* b, const Csr* c, const u64* z,
              const u64* r, const u64* s, u64* proof, int threads, double* t_ms) {
  init_all();
  DISPATCH(curve, { return C::prove(pk, ni, nc, nw, a, b, c, z, r, s, proof, threads, t_ms); })
}
}
