// msm_ba.cuh -- batched-affine pre-reduction of the sorted bucket entries (EXPERIMENTAL, off by default: G16_MSM_BA).
//
// The bucket accumulation of msm.cuh spends 10 field multiplications per entry (XYZZ mixed addition).  An affine
// addition costs 3 once the inverse of (x2 - x1) is known, and Montgomery's trick shares one inversion between any
// number of independent additions at 3 more multiplications each.  Inside one bucket the entries can be summed as a
// binary tree, and every level of that tree is a set of independent additions:
//
//   round r:  list_r (points, grouped by bucket, off_r[b] = start of bucket b)  ->  list_{r+1},
//             count_{r+1}[b] = ceil(count_r[b] / 2);   output p of bucket b = in[2p] + in[2p+1]  (or in[2p] alone).
//
// Round 0 reads its points through the sorted index list (base index, sign bit); later rounds read the previous list.
// After R rounds 1 - 2^-R of all additions are done and the remaining list goes through msm_accum_l0 unchanged.
// One round is three launches:
//   forward : thread t owns outputs j = k*T + t (k < m): d_j = x2 - x1, exclusive running product -> pre[j],
//             thread product -> prod[t]
//   combine : lane g owns G thread products: Montgomery's trick over them, ONE inversion per lane, prod[t] <- prod[t]^-1
//   backward: thread t walks its outputs in reverse, recovers 1/d_j, finishes the additions, writes list_{r+1}
// = 6 multiplications per addition + (3 + inv/G)/m for the combine (inv ~ 570 multiplications with Fermat; the
// safegcd inversion of fp_inv.cuh, ~10x fewer instructions, is selectable: G16_BA_INV_GCD=1, not yet run on a GPU).
//
// Every per-thread body below is __host__ __device__ and free of warp intrinsics, so tests/host/ba_check.cu runs
// the same code on the CPU against plain XYZZ sums (exceptional cases included: equal points, opposite points,
// identities in the lists).
#pragma once
#include "ec.cuh"
#include "fp_inv.cuh"

namespace g16 {

template <class F>
struct BaRound {
  const Affine<F>* in;       // round 0: base table (gathered through sidx); later: the previous list
  const uint32_t* sidx;      // round 0 only: (base index | sign << 31) per sorted entry; nullptr afterwards
  const uint32_t* off_in;    // [nkeys + 1] bucket offsets of the input list
  const uint32_t* off_out;   // [nkeys + 1] bucket offsets of the output list
  uint32_t nkeys;
  uint32_t m;                // outputs per thread
  uint32_t G;                // thread products per combine lane
  uint32_t inv_gcd;          // combine: 1 = safegcd inversion (fp_inv.cuh), 0 = Fermat
  F* pre;                    // [outputs]  running product of the thread before output j
  uint32_t* key;             // [outputs]  bucket of output j
  uint32_t* ident;           // [outputs]  j  (the index list msm_accum_l0 wants for the final list)
  F* prod;                   // [threads]  thread product, then its inverse
  F* pre2;                   // [threads]  running product of the lane before thread t
  Affine<F>* out;            // [outputs]
};

G16_HD uint64_t ba_threads(uint64_t outputs, uint32_t m) { return (outputs + m - 1) / m; }
// upper bound of the list length after one more round
inline uint64_t ba_next_max(uint64_t len_max, uint32_t nkeys) { return (len_max + nkeys + 1) / 2; }

template <class F>
G16_HD F ba_ld(const F* p) {
#ifdef __CUDA_ARCH__
  F v;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&v);
#pragma unroll
  for (int j = 0; j < (int)(sizeof(F) / 16); j++) d[j] = __ldg(s + j);
  return v;
#else
  return *p;
#endif
}

// largest b with off[b] <= j   (j < off[nkeys], so off[b + 1] > j: empty buckets are skipped)
G16_HD uint32_t ba_bucket_of(const uint32_t* off, uint32_t nkeys, uint32_t j) {
  uint32_t lo = 0, hi = nkeys - 1;
  while (lo < hi) {
    const uint32_t mid = (lo + hi + 1) >> 1;
    if (off[mid] <= j) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// where input i of the round lives, and whether its y is to be negated
template <class F>
struct BaSrc {
  const Affine<F>* p;
  bool neg;
};
template <class F>
G16_HD BaSrc<F> ba_src(const BaRound<F>& a, uint32_t i) {
  if (a.sidx) {
    const uint32_t ix = a.sidx[i];
    return {a.in + (ix & 0x7fffffffu), (ix >> 31) != 0};
  }
  return {a.in + i, false};
}
template <class F>
G16_HD F ba_y(const BaSrc<F>& s) {
  F y = ba_ld(&s.p->y);
  return s.neg ? F::neg(y) : y;
}

// kind of a pair: 0 chord (d = x2 - x1), 1 tangent (d = 2 y1), 2 result = P1, 3 result = P2, 4 result = identity
enum { BA_CHORD = 0, BA_TANGENT = 1, BA_FIRST = 2, BA_SECOND = 3, BA_INF = 4 };
template <class F>
G16_HD int ba_classify(const F& x1, const F& y1, const F& x2, const F& y2, F& d) {
  const bool inf1 = x1.is_zero() && y1.is_zero(), inf2 = x2.is_zero() && y2.is_zero();
  if (inf1) return BA_SECOND;
  if (inf2) return BA_FIRST;
  if (x1 == x2) {
    if (y1 == y2 && !y1.is_zero()) { d = F::dbl(y1); return BA_TANGENT; }
    return BA_INF;
  }
  d = F::sub(x2, x1);
  return BA_CHORD;
}

template <class F>
G16_HD void ba_forward(const BaRound<F>& a, uint64_t t) {
  const uint32_t M = a.off_out[a.nkeys];
  const uint64_t T = ba_threads(M, a.m);
  if (t >= T) return;
  F run = F::one();
  for (uint32_t k = 0; k < a.m; k++) {
    const uint64_t j64 = (uint64_t)k * T + t;
    if (j64 >= M) break;
    const uint32_t j = (uint32_t)j64;
    const uint32_t b = ba_bucket_of(a.off_out, a.nkeys, j);
    a.key[j] = b;
    const uint32_t p = j - a.off_out[b], i0 = a.off_in[b] + 2 * p, cnt = a.off_in[b + 1] - a.off_in[b];
    if (2 * p + 1 >= cnt) continue;   // odd one out: copied by the backward pass
    const BaSrc<F> s1 = ba_src(a, i0), s2 = ba_src(a, i0 + 1);
    const F x1 = ba_ld(&s1.p->x), x2 = ba_ld(&s2.p->x);
    F d;
    int kind = BA_CHORD;
    if (x1.is_zero() || x2.is_zero() || x1 == x2) kind = ba_classify(x1, ba_y(s1), x2, ba_y(s2), d);
    else d = F::sub(x2, x1);
    a.pre[j] = run;
    if (kind <= BA_TANGENT) run = F::mul(run, d);
  }
  a.prod[t] = run;
}

// the one inversion of a combine lane
template <class P>
G16_HD Fp<P> ba_inv(const Fp<P>& a, bool gcd) { return gcd ? fp_inv_safegcd<P>(a) : Fp<P>::inv(a); }
template <class P, int NR>
G16_HD Fp2<P, NR> ba_inv(const Fp2<P, NR>& a, bool gcd) {
  using B = Fp<P>;
  const B n = B::add(B::sqr(a.c0), Fp2<P, NR>::mul_nr(B::sqr(a.c1)));
  const B ni = ba_inv(n, gcd);
  return {B::mul(a.c0, ni), B::neg(B::mul(a.c1, ni))};
}

template <class F>
G16_HD void ba_combine(const BaRound<F>& a, uint64_t g) {
  const uint32_t M = a.off_out[a.nkeys];
  const uint64_t T = ba_threads(M, a.m);
  const uint64_t lo = g * a.G;
  if (lo >= T) return;
  const uint64_t hi = (lo + a.G < T) ? lo + a.G : T;
  F acc = F::one();
  for (uint64_t k = lo; k < hi; k++) {
    a.pre2[k] = acc;
    acc = F::mul(acc, a.prod[k]);
  }
  F inv = ba_inv(acc, a.inv_gcd != 0);   // never zero: every factor is x2 - x1 != 0, 2 y1 != 0 or one
  for (uint64_t k = hi; k-- > lo;) {
    const F pk = a.prod[k];
    a.prod[k] = F::mul(inv, a.pre2[k]);
    inv = F::mul(inv, pk);
  }
}

template <class F>
G16_HD void ba_backward(const BaRound<F>& a, uint64_t t) {
  const uint32_t M = a.off_out[a.nkeys];
  const uint64_t T = ba_threads(M, a.m);
  if (t >= T) return;
  F run_inv = a.prod[t];
  uint32_t kn = 0;   // outputs of this thread
  while (kn < a.m && (uint64_t)kn * T + t < M) kn++;
  for (uint32_t k = kn; k-- > 0;) {
    const uint32_t j = (uint32_t)((uint64_t)k * T + t);
    const uint32_t b = a.key[j];
    const uint32_t p = j - a.off_out[b], i0 = a.off_in[b] + 2 * p, cnt = a.off_in[b + 1] - a.off_in[b];
    const BaSrc<F> s1 = ba_src(a, i0);
    Affine<F> r{ba_ld(&s1.p->x), ba_y(s1)};
    if (2 * p + 1 < cnt) {
      const BaSrc<F> s2 = ba_src(a, i0 + 1);
      const F x2 = ba_ld(&s2.p->x), y2 = ba_y(s2);
      F d;
      const int kind = ba_classify(r.x, r.y, x2, y2, d);
      if (kind <= BA_TANGENT) {
        const F inv_d = F::mul(run_inv, a.pre[j]);
        run_inv = F::mul(run_inv, d);
        F lam, x3;
        if (kind == BA_CHORD) {
          lam = F::mul(F::sub(y2, r.y), inv_d);
          x3 = F::sub(F::sub(F::sqr(lam), r.x), x2);
        } else {
          const F xx = F::sqr(r.x);
          lam = F::mul(F::add(F::dbl(xx), xx), inv_d);
          x3 = F::sub(F::sqr(lam), F::dbl(r.x));
        }
        r.y = F::sub(F::mul(lam, F::sub(r.x, x3)), r.y);
        r.x = x3;
      } else if (kind == BA_SECOND) {
        r.x = x2;
        r.y = y2;
      } else if (kind == BA_INF) {
        r = Affine<F>::inf();
      }
    }
    a.out[j] = r;
    a.ident[j] = j;
  }
}

// Same pass with the operands consumed as early as possible (x2 and y2 are folded into x1 + x2, x2 - x1 and y2 - y1
// and dropped before the multiplications start): fewer live limbs for the Fq2 instantiation, whose straightforward
// version above needs more than 255 registers.  Selected with G16_BA_LEAN=1 until measured on a GPU.
template <class F>
G16_HD void ba_backward_lean(const BaRound<F>& a, uint64_t t) {
  const uint32_t M = a.off_out[a.nkeys];
  const uint64_t T = ba_threads(M, a.m);
  if (t >= T) return;
  F run_inv = a.prod[t];
  uint32_t kn = 0;
  while (kn < a.m && (uint64_t)kn * T + t < M) kn++;
  for (uint32_t k = kn; k-- > 0;) {
    const uint32_t j = (uint32_t)((uint64_t)k * T + t);
    const uint32_t b = a.key[j];
    const uint32_t p = j - a.off_out[b], i0 = a.off_in[b] + 2 * p, cnt = a.off_in[b + 1] - a.off_in[b];
    const BaSrc<F> s1 = ba_src(a, i0);
    Affine<F> r;
    r.x = ba_ld(&s1.p->x);
    if (2 * p + 1 >= cnt) {          // odd one out: copy
      r.y = ba_y(s1);
    } else {
      const BaSrc<F> s2 = ba_src(a, i0 + 1);
      F sx, d;                        // x1 + x2 and the denominator
      int kind = BA_CHORD;
      {
        const F x2 = ba_ld(&s2.p->x);
        if (r.x.is_zero() || x2.is_zero() || r.x == x2) kind = ba_classify(r.x, ba_y(s1), x2, ba_y(s2), d);
        else d = F::sub(x2, r.x);
        sx = F::add(r.x, x2);
      }
      if (kind <= BA_TANGENT) {
        F lam = F::mul(run_inv, a.pre[j]);           // 1 / d
        run_inv = F::mul(run_inv, d);
        r.y = ba_y(s1);
        if (kind == BA_CHORD) {
          lam = F::mul(F::sub(ba_y(s2), r.y), lam);
        } else {
          const F xx = F::sqr(r.x);
          lam = F::mul(F::add(F::dbl(xx), xx), lam);
        }
        const F x3 = F::sub(F::sqr(lam), sx);        // tangent: sx = 2 x1
        r.y = F::sub(F::mul(lam, F::sub(r.x, x3)), r.y);
        r.x = x3;
      } else if (kind == BA_FIRST) {
        r.y = ba_y(s1);
      } else if (kind == BA_SECOND) {
        r.x = ba_ld(&s2.p->x);
        r.y = ba_y(s2);
      } else {
        r = Affine<F>::inf();
      }
    }
    a.out[j] = r;
    a.ident[j] = j;
  }
}

#ifdef __CUDACC__
template <class F>
__global__ void __launch_bounds__(128) ba_backward_lean_kernel(BaRound<F> a) {
  ba_backward_lean<F>(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
template <class F>
__global__ void __launch_bounds__(128) ba_forward_kernel(BaRound<F> a) {
  ba_forward<F>(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
template <class F>
__global__ void __launch_bounds__(32) ba_combine_kernel(BaRound<F> a) {
  ba_combine<F>(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
template <class F>
__global__ void __launch_bounds__(128) ba_backward_kernel(BaRound<F> a) {
  ba_backward<F>(a, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

// Bucket offsets of every round from the offsets of the sorted entries, one block: off_all[r][b], r = 0..R.
// Thread i scans the contiguous key range [i * per, (i + 1) * per).  (block_exclusive_scan_1024: msm.cuh, which includes this file.)
static __global__ void __launch_bounds__(1024) ba_offsets_kernel(const uint32_t* __restrict__ off0, uint32_t nkeys, int R,
                                                                 uint32_t* __restrict__ off_all) {
  __shared__ uint32_t sh[64];
  const uint32_t per = (nkeys + 1023) / 1024;
  const uint32_t lo = min(nkeys, threadIdx.x * per), hi = min(nkeys, lo + per);
  for (uint32_t b = threadIdx.x; b <= nkeys; b += 1024) off_all[b] = off0[b];
  __syncthreads();
  for (int r = 1; r <= R; r++) {
    const uint32_t* src = off_all + (size_t)(r - 1) * (nkeys + 1);
    uint32_t* dst = off_all + (size_t)r * (nkeys + 1);
    uint32_t s = 0;
    for (uint32_t b = lo; b < hi; b++) s += (src[b + 1] - src[b] + 1) >> 1;
    uint32_t tot;
    uint32_t run = block_exclusive_scan_1024(s, sh, &tot);
    for (uint32_t b = lo; b < hi; b++) { dst[b] = run; run += (src[b + 1] - src[b] + 1) >> 1; }
    if (threadIdx.x == 0) dst[nkeys] = tot;
    __syncthreads();
  }
}
#endif  // __CUDACC__

}  // namespace g16
