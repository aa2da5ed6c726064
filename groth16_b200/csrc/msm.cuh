// msm.cuh -- variable-base multi-scalar multiplication on sm_100a (Pippenger, signed digits, sort-by-bucket).
//
// Replaces ark-ec 0.5.0 `VariableBaseMSM::msm_bigint` as called at /root/reference/src/prover.rs:66 (H query),
// :74 (L query) and :262 (A, B-in-G1, B-in-G2 via calculate_coeff).  The group element returned is identical to
// the reference's (EC addition is exactly associative/commutative), whatever the window size or summation order.
//
// Pipeline (all on one stream, no host synchronisation until the W window sums are read back):
//   1. msm_digits<COUNT>   scalar -> W signed c-bit digits; histogram of (window, |digit|) keys (warp-aggregated atomics)
//   2. msm_scan            exclusive prefix sum of the histogram -> bucket offsets
//   3. msm_digits<SCATTER> counting-sort scatter: sorted (base index | sign) and key per entry
//   4. msm_accum_l0        load-balanced segmented reduction: every thread owns K0 consecutive sorted entries,
//                          mixed-adds them (XYZZ += affine, gathered from the resident base array), writes buckets
//                          that are complete inside its chunk and emits <= 2 boundary partials
//   5. msm_accum_ln        the same reduction over the partial list, level by level, until one thread remains
//   6. msm_bucket_reduce   per (window, segment of L buckets): running-sum trick + (segment offset) * sum
//   7. msm_sum_groups      tree-sum of the segment results -> one point per window
//   host: Horner over windows (c doublings each), prover.rs semantics preserved.
// Skewed scalar distributions (boolean witnesses, the reference's DummyCircuit whose witness is constant,
// benches/bench.rs:43-54) cost the same as uniform ones: work is split by sorted position, not by bucket.
#pragma once
#include <cuda_runtime.h>
#include "ec.cuh"

namespace g16 {

static constexpr uint32_t MSM_INVALID = 0xffffffffu;

struct MsmGeom {
  uint32_t n;        // number of (scalar, base) pairs
  int c;             // window bits
  int W;             // number of c-bit windows of a scalar
  int ne;            // effective windows: window w = j*ne + e lands in bucket set e and uses base copy j
  int copies;        // ceil(W / ne) precomputed multiples 2^(c*ne*j) * P of every base (1 = no precomputation)
  uint32_t B;        // buckets per effective window = 2^(c-1)  (digit magnitudes 1..B)
  uint32_t nkeys;    // ne * B
  uint64_t max_entries;  // n * W
};

inline int msm_pick_c(uint64_t n) {
  int lg = 0;
  while ((1ull << lg) < n) lg++;
  int c = lg - 4;
  if (c < 3) c = 3;
  if (c > 16) c = 16;
  return c;
}
// ne_req <= 0: no precomputation (ne = W).  Otherwise the requested number of effective windows (1 = every window
// of a scalar shares one bucket set, which needs W precomputed multiples per base).
inline MsmGeom msm_geom(uint64_t n, int scalar_bits, int c_override = 0, int ne_req = 0) {
  MsmGeom g;
  g.n = (uint32_t)n;
  g.c = c_override > 0 ? c_override : msm_pick_c(n);
  g.W = (scalar_bits + 1 + g.c - 1) / g.c;   // +1: room for the top signed-digit carry
  g.ne = (ne_req <= 0 || ne_req > g.W) ? g.W : ne_req;
  g.copies = (g.W + g.ne - 1) / g.ne;
  g.B = 1u << (g.c - 1);
  g.nkeys = (uint32_t)g.ne * g.B;
  g.max_entries = (uint64_t)n * g.W;
  return g;
}

// ------------------------------------------------------------------------------------------------
// 1/3. digit extraction + histogram / scatter
// ------------------------------------------------------------------------------------------------
// Signed-digit recoding of a canonical scalar k < 2^bits:  k = sum_w d_w 2^(c w), d_w in [-2^(c-1), 2^(c-1)].
template <class FrF, bool SCATTER>
__global__ void __launch_bounds__(256) msm_digits(const uint32_t* __restrict__ scalars, int scalars_mont,
                                                  const uint8_t* __restrict__ skip, MsmGeom g,
                                                  uint32_t* __restrict__ counters,  // COUNT: histogram; SCATTER: cursors
                                                  uint32_t* __restrict__ sidx, uint32_t* __restrict__ skey) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  bool live = i < g.n;
  FrF s = FrF::zero();
  if (live) {
    const uint4* p = reinterpret_cast<const uint4*>(scalars) + (size_t)i * 2;
    uint4 lo = __ldg(p), hi = __ldg(p + 1);
    s.v[0] = lo.x; s.v[1] = lo.y; s.v[2] = lo.z; s.v[3] = lo.w;
    s.v[4] = hi.x; s.v[5] = hi.y; s.v[6] = hi.z; s.v[7] = hi.w;
    if (skip && skip[i]) live = false;
  }
  if (scalars_mont) s = FrF::from_mont(s);   // `into_bigint`, prover.rs:64,71,82
  uint32_t carry = 0;
  for (int w = 0; w < g.W; w++) {
    // c <= 16 bits starting at bit w*c
    const int bit = w * g.c;
    const int limb = bit >> 5, sh = bit & 31;
    uint32_t raw = 0;
    if (limb < 8) {
      uint64_t two = s.v[limb];
      if (limb + 1 < 8) two |= (uint64_t)s.v[limb + 1] << 32;
      raw = (uint32_t)(two >> sh) & ((1u << g.c) - 1);
    }
    raw += carry;
    uint32_t neg = 0;
    carry = 0;
    if (raw > g.B) { raw = (1u << g.c) - raw; neg = 1; carry = 1; }
    const bool emit = live && raw != 0;
    const int e = w % g.ne, j = w / g.ne;
    const uint32_t key = emit ? (uint32_t)e * g.B + (raw - 1) : MSM_INVALID;
    // warp-aggregated atomic: one atomicAdd per distinct key in the warp
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    if (emit) {
      const uint32_t rank = __popc(peers & ((1u << lane) - 1));
      const int leader = __ffs(peers) - 1;
      uint32_t base = 0;
      if ((int)lane == leader) base = atomicAdd(&counters[key], (uint32_t)__popc(peers));
      base = __shfl_sync(peers, base, leader);
      if (SCATTER) {
        const uint32_t pos = base + rank;
        sidx[pos] = ((uint32_t)j * g.n + i) | (neg << 31);
        skey[pos] = key;
      }
    }
  }
}

// 2. exclusive scan in three small launches: per-block scan (4096 keys per block), scan of the block totals, fix-up.
// `cursors` may alias `hist` (the histogram is turned into the scatter cursors in place).  offsets[nkeys] = total.
static constexpr int SCAN_ITEMS = 4;
static constexpr int SCAN_BLOCK = 1024 * SCAN_ITEMS;
static __device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t* sh, uint32_t* total) {
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= (uint32_t)d) x += y;
  }
  if (lane == 31) sh[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = sh[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= (uint32_t)d) w += y;
    }
    sh[32 + lane] = w;
  }
  __syncthreads();
  const uint32_t warp_off = wid ? sh[32 + wid - 1] : 0;
  *total = sh[63];
  return warp_off + x - v;
}
static __global__ void __launch_bounds__(1024) msm_scan_blocks(const uint32_t* hist, uint32_t nkeys, uint32_t* offsets,
                                                               uint32_t* block_tot) {
  __shared__ uint32_t sh[64];
  const uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < nkeys) ? hist[base + k] : 0; s += v[k]; }
  uint32_t tot;
  uint32_t run = block_exclusive_scan_1024(s, sh, &tot);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < nkeys) offsets[base + k] = run; run += v[k]; }
  if (threadIdx.x == 0) block_tot[blockIdx.x] = tot;
}
static __global__ void __launch_bounds__(1024) msm_scan_tops(uint32_t* block_tot, uint32_t nblocks, uint32_t* total_out) {
  __shared__ uint32_t sh[64];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < nblocks; b0 += 1024) {
    const uint32_t i = b0 + threadIdx.x;
    const uint32_t v = i < nblocks ? block_tot[i] : 0;
    uint32_t tot;
    const uint32_t ex = block_exclusive_scan_1024(v, sh, &tot);
    const uint32_t c = carry;
    __syncthreads();
    if (i < nblocks) block_tot[i] = c + ex;
    if (threadIdx.x == 0) carry = c + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}
static __global__ void __launch_bounds__(1024) msm_scan_fix(uint32_t* offsets, uint32_t nkeys, const uint32_t* block_tot,
                                                            uint32_t* cursors) {
  const uint32_t add = block_tot[blockIdx.x];
  const uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++)
    if (base + k < nkeys) {
      const uint32_t o = offsets[base + k] + add;
      offsets[base + k] = o;
      cursors[base + k] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// 4/5. load-balanced segmented bucket accumulation
// ------------------------------------------------------------------------------------------------
// Every thread t of a level owns two output slots (2t: "head" partial, 2t+1: "tail" partial) and always writes
// both keys (MSM_INVALID when unused), so no memset of the partial lists is needed.
template <class F>
struct MsmEmit {
  XYZZ<F>* buckets;
  uint32_t* okeys;
  XYZZ<F>* opts;
  uint64_t t;
  bool wrote_head, wrote_tail;
  __device__ __forceinline__ void flush(uint32_t key, const XYZZ<F>& acc, bool head, bool tail) {
    if (!head && !tail) {
      buckets[key] = acc;   // the whole bucket was summed here: written exactly once, no atomics
    } else {
      const uint64_t slot = 2 * t + (head ? 0 : 1);
      okeys[slot] = key;
      opts[slot] = acc;
      if (head) wrote_head = true; else wrote_tail = true;
    }
  }
  __device__ __forceinline__ void finish() {
    if (!wrote_head) okeys[2 * t] = MSM_INVALID;
    if (!wrote_tail) okeys[2 * t + 1] = MSM_INVALID;
  }
};

template <class F>
__device__ __forceinline__ Affine<F> load_affine(const Affine<F>* __restrict__ bases, uint32_t idx) {
  Affine<F> p;
  constexpr int NV = sizeof(Affine<F>) / 16;
  const uint4* src = reinterpret_cast<const uint4*>(bases + idx);
  uint4* dst = reinterpret_cast<uint4*>(&p);
#pragma unroll
  for (int j = 0; j < NV; j++) dst[j] = __ldg(src + j);
  return p;
}

// Level 0: grid covers T0 = ceil(max_entries / K0) threads; threads past the real entry count only clear their slots.
template <class F, int K0>
__global__ void __launch_bounds__(128) msm_accum_l0(const Affine<F>* __restrict__ bases,
                                                    const uint32_t* __restrict__ sidx,
                                                    const uint32_t* __restrict__ skey,
                                                    const uint32_t* __restrict__ total_ptr, uint64_t T0,
                                                    XYZZ<F>* __restrict__ buckets, uint32_t* __restrict__ okeys,
                                                    XYZZ<F>* __restrict__ opts) {
  const uint32_t M = *total_ptr;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T0) return;
  MsmEmit<F> em{buckets, okeys, opts, t, false, false};
  const uint64_t begin = t * K0;
  if (begin >= M) { em.finish(); return; }
  const uint32_t end = (uint32_t)min((uint64_t)M, begin + K0);
  const uint32_t prev = begin > 0 ? skey[begin - 1] : MSM_INVALID;
  const uint32_t next = end < M ? skey[end] : MSM_INVALID;
  uint32_t cur = skey[begin];
  bool first_seg = true;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t e = (uint32_t)begin; e < end; e++) {
    const uint32_t k = skey[e];
    if (k != cur) {
      em.flush(cur, acc, first_seg && prev == cur, false);
      first_seg = false;
      cur = k;
      acc = XYZZ<F>::inf();
    }
    const uint32_t ix = sidx[e];
    const Affine<F> p = load_affine(bases, ix & 0x7fffffffu);
    acc.madd_inline(p, (ix >> 31) != 0);
  }
  em.flush(cur, acc, first_seg && prev == cur, next == cur);
  em.finish();
}

// One level >= 1 for thread t over slots [t*KF, t*KF + KF) of a partial list of S slots.  Keys may contain
// MSM_INVALID holes; inside a run of equal keys there is at most one hole between neighbours, so a
// look-back / look-ahead of two slots decides whether a run continues across the chunk boundary.
template <class F, int KF>
__device__ __forceinline__ void msm_level_step(const uint32_t* ikeys, const XYZZ<F>* ipts, uint64_t S, uint64_t t,
                                               XYZZ<F>* buckets, uint32_t* okeys, XYZZ<F>* opts) {
  MsmEmit<F> em{buckets, okeys, opts, t, false, false};
  const uint64_t begin = t * KF;
  if (begin >= S) { em.finish(); return; }
  const uint64_t end = min(S, begin + KF);
  uint32_t prev = MSM_INVALID, next = MSM_INVALID;
  if (begin >= 1) prev = ikeys[begin - 1];
  if (prev == MSM_INVALID && begin >= 2) prev = ikeys[begin - 2];
  if (end < S) next = ikeys[end];
  if (next == MSM_INVALID && end + 1 < S) next = ikeys[end + 1];
  uint32_t cur = MSM_INVALID;
  bool have = false, first_seg = true;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint64_t e = begin; e < end; e++) {
    const uint32_t k = ikeys[e];
    if (k == MSM_INVALID) continue;
    if (!have) {
      have = true;
      cur = k;
      acc = ipts[e];
    } else if (k != cur) {
      em.flush(cur, acc, first_seg && prev == cur, false);
      first_seg = false;
      cur = k;
      acc = ipts[e];
    } else {
      acc.add(ipts[e]);
    }
  }
  if (have) em.flush(cur, acc, first_seg && prev == cur, next == cur);
  em.finish();
}
template <class F, int KF>
__global__ void __launch_bounds__(128) msm_accum_ln(const uint32_t* __restrict__ ikeys,
                                                    const XYZZ<F>* __restrict__ ipts, uint64_t S, uint64_t T,
                                                    XYZZ<F>* __restrict__ buckets, uint32_t* __restrict__ okeys,
                                                    XYZZ<F>* __restrict__ opts) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  msm_level_step<F, KF>(ikeys, ipts, S, t, buckets, okeys, opts);
}
// All remaining levels in one block once the partial list is short (S <= 2 * blockDim * KF): no launch gaps.
template <class F, int KF>
__global__ void __launch_bounds__(256) msm_accum_tail(uint32_t* k0, XYZZ<F>* p0, uint32_t* k1, XYZZ<F>* p1, uint64_t S,
                                                      XYZZ<F>* buckets) {
  uint32_t *ik = k0, *ok = k1;
  XYZZ<F>*ip = p0, *op = p1;
  while (true) {
    const uint64_t T = (S + KF - 1) / KF;
    for (uint64_t t = threadIdx.x; t < T; t += blockDim.x) msm_level_step<F, KF>(ik, ip, S, t, buckets, ok, op);
    __syncthreads();
    if (T == 1) break;
    S = 2 * T;
    uint32_t* tk = ik; ik = ok; ok = tk;
    XYZZ<F>* tp = ip; ip = op; op = tp;
  }
}

// ------------------------------------------------------------------------------------------------
// 6/7. bucket reduction
// ------------------------------------------------------------------------------------------------
// out[w*nseg + seg] = sum_{j<L} (seg*L + j + 1) * bucket[w*B + seg*L + j]
template <class F>
__global__ void __launch_bounds__(128) msm_bucket_reduce(const XYZZ<F>* __restrict__ buckets, uint32_t B, uint32_t L,
                                                         uint32_t total_segs, XYZZ<F>* __restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total_segs) return;
  const uint32_t nseg = B / L;
  const uint32_t w = t / nseg, seg = t % nseg;
  const XYZZ<F>* base = buckets + (size_t)w * B + (size_t)seg * L;
  XYZZ<F> running = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
  for (int j = (int)L - 1; j >= 0; j--) {
    running.add(base[j]);
    acc.add(running);
  }
  if (seg) {
    uint32_t k = seg * L;
    acc.add(running.mul_u32(&k, 1));
  }
  out[t] = acc;
}

// out[g] = sum_{j<R} in[g*R + j]   (g < ngroups)
template <class F>
__global__ void __launch_bounds__(128) msm_sum_groups(const XYZZ<F>* __restrict__ in, uint32_t R, uint32_t ngroups,
                                                      XYZZ<F>* __restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ngroups) return;
  XYZZ<F> acc = in[(size_t)t * R];
  for (uint32_t j = 1; j < R; j++) acc.add(in[(size_t)t * R + j]);
  out[t] = acc;
}

// infinity mask of a base array (x == y == 0), computed once when a query is made resident
template <class F>
__global__ void msm_inf_mask(const Affine<F>* __restrict__ bases, uint32_t n, uint8_t* __restrict__ mask) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Affine<F> p = load_affine(bases, i);
  mask[i] = p.is_inf() ? 1 : 0;
}

// Base precomputation: out[j*n + i] = 2^(shift*j) * in[i] for j < copies, as affine points (one inversion per base:
// the copies of a base are normalised together with Montgomery's trick).  `out` copy 0 may alias `in`.
static constexpr int MSM_MAX_COPIES = 20;
template <class F>
__global__ void __launch_bounds__(128) msm_precompute(const Affine<F>* in, uint32_t n, int copies, int shift,
                                                      Affine<F>* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Affine<F> p = load_affine(in, i);
  out[i] = p;
  XYZZ<F> q = XYZZ<F>::from_affine(p);
  F Xs[MSM_MAX_COPIES], Ys[MSM_MAX_COPIES], ZZs[MSM_MAX_COPIES], ZZZs[MSM_MAX_COPIES], pref[MSM_MAX_COPIES];
  F acc = F::one();
  int live = 0;   // copies 1..live are finite
  for (int j = 1; j < copies; j++) {
    for (int s = 0; s < shift; s++) q.dbl_inplace();
    if (q.is_inf()) break;
    Xs[j] = q.X; Ys[j] = q.Y; ZZs[j] = q.ZZ; ZZZs[j] = q.ZZZ;
    pref[j] = acc;
    acc = F::mul(acc, q.ZZZ);
    live = j;
  }
  for (int j = live + 1; j < copies; j++) out[(size_t)j * n + i] = Affine<F>::inf();
  if (live == 0) return;
  F inv = F::inv(acc);
  for (int j = live; j >= 1; j--) {
    const F zi = F::mul(inv, pref[j]);   // 1 / ZZZ_j
    inv = F::mul(inv, ZZZs[j]);
    const F z = F::mul(zi, ZZs[j]);      // 1 / Z_j
    const F zi2 = F::sqr(z);
    Affine<F> a{F::mul(Xs[j], zi2), F::mul(Ys[j], zi)};
    out[(size_t)j * n + i] = a;
  }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

static constexpr int MSM_K0 = 64;      // sorted entries per thread, level 0
static constexpr int MSM_KF = 4;       // partial slots per thread, levels >= 1
static constexpr int MSM_TAIL_S = 2048;  // partial-list length at which the remaining levels fuse into one block
static constexpr int MSM_GRP = 8;      // fan-in of the window-sum tree

inline uint32_t msm_seg_len(const MsmGeom& g) {
  // buckets per thread in the bucket reduction: keep >= ~8k threads in flight, at most 16 buckets per thread
  uint32_t L = 16;
  while (L > 2 && (uint64_t)g.ne * (g.B / L) < 8192) L >>= 1;
  if (L > g.B) L = g.B;
  return L;
}

template <class F>
struct MsmWorkspace {
  DevBuf counters, offsets, blocktot, sidx, skey, buckets, pk0, pp0, pk1, pp1, seg0, seg1;
  XYZZ<F>* h_winsums = nullptr;  // pinned host staging for the window sums
  int h_cap = 0;
  cudaError_t prepare(const MsmGeom& g) {
    cudaError_t e;
    const uint64_t T0 = (g.max_entries + MSM_K0 - 1) / MSM_K0;
    const uint64_t S1 = 2 * T0;
    const uint64_t T1 = (S1 + MSM_KF - 1) / MSM_KF;
    const uint64_t S2 = 2 * T1;
#define G16_TRY(x) if ((e = (x)) != cudaSuccess) return e
    G16_TRY(counters.reserve((size_t)(g.nkeys + 1) * 4));
    G16_TRY(offsets.reserve((size_t)(g.nkeys + 1) * 4));
    G16_TRY(blocktot.reserve((size_t)((g.nkeys + SCAN_BLOCK - 1) / SCAN_BLOCK + 1) * 4));
    G16_TRY(sidx.reserve(g.max_entries * 4 + 16));
    G16_TRY(skey.reserve(g.max_entries * 4 + 16));
    G16_TRY(buckets.reserve((size_t)g.nkeys * sizeof(XYZZ<F>)));
    G16_TRY(pk0.reserve(S1 * 4 + 16));
    G16_TRY(pp0.reserve(S1 * sizeof(XYZZ<F>)));
    G16_TRY(pk1.reserve(S2 * 4 + 16));
    G16_TRY(pp1.reserve(S2 * sizeof(XYZZ<F>)));
    const uint32_t L = msm_seg_len(g);
    const uint64_t nsegs = (uint64_t)g.ne * (g.B / L);
    G16_TRY(seg0.reserve(nsegs * sizeof(XYZZ<F>)));
    G16_TRY(seg1.reserve((nsegs / 2 + g.ne) * sizeof(XYZZ<F>)));
    if (h_cap < g.ne) {
      if (h_winsums) cudaFreeHost(h_winsums);
      h_winsums = nullptr;
      G16_TRY(cudaMallocHost(&h_winsums, (size_t)g.ne * sizeof(XYZZ<F>)));
      h_cap = g.ne;
    }
#undef G16_TRY
    return cudaSuccess;
  }
  void release() {
    counters.release(); offsets.release(); blocktot.release(); sidx.release(); skey.release(); buckets.release();
    pk0.release(); pp0.release(); pk1.release(); pp1.release(); seg0.release(); seg1.release();
    if (h_winsums) cudaFreeHost(h_winsums);
    h_winsums = nullptr;
    h_cap = 0;
  }
};

struct MsmCounters {  // launch bookkeeping for bench.py's gpu_launches
  unsigned long long launches = 0;
};

// Enqueue one MSM on `st`.  d_bases holds g.copies * g.n affine points (copy-major); d_scalars / d_skip are device
// pointers; the g.ne window sums land in ws.h_winsums once the stream is synchronised (msm_finish: host Horner).
template <class F, class FrF>
cudaError_t msm_enqueue(cudaStream_t st, MsmWorkspace<F>& ws, const MsmGeom& g, const Affine<F>* d_bases,
                        const uint8_t* d_skip, const uint32_t* d_scalars, bool scalars_mont, MsmCounters* ctr,
                        cudaEvent_t ev_acc0 = nullptr, cudaEvent_t ev_acc1 = nullptr) {
  cudaError_t e;
  if (g.n == 0) return cudaSuccess;
  if ((e = ws.prepare(g)) != cudaSuccess) return e;
  uint32_t* counters = ws.counters.template as<uint32_t>();
  uint32_t* offsets = ws.offsets.template as<uint32_t>();
  uint32_t* blocktot = ws.blocktot.template as<uint32_t>();
  uint32_t* sidx = ws.sidx.template as<uint32_t>();
  uint32_t* skey = ws.skey.template as<uint32_t>();
  XYZZ<F>* buckets = ws.buckets.template as<XYZZ<F>>();
  unsigned long long nl = 0;
  cudaMemsetAsync(counters, 0, (size_t)(g.nkeys + 1) * 4, st);
  cudaMemsetAsync(buckets, 0, (size_t)g.nkeys * sizeof(XYZZ<F>), st);
  const uint32_t nb = (g.n + 255) / 256;
  msm_digits<FrF, false><<<nb, 256, 0, st>>>(d_scalars, scalars_mont ? 1 : 0, d_skip, g, counters, nullptr, nullptr);
  const uint32_t sb = (g.nkeys + SCAN_BLOCK - 1) / SCAN_BLOCK;
  msm_scan_blocks<<<sb, 1024, 0, st>>>(counters, g.nkeys, offsets, blocktot);
  msm_scan_tops<<<1, 1024, 0, st>>>(blocktot, sb, offsets + g.nkeys);
  msm_scan_fix<<<sb, 1024, 0, st>>>(offsets, g.nkeys, blocktot, counters);
  msm_digits<FrF, true><<<nb, 256, 0, st>>>(d_scalars, scalars_mont ? 1 : 0, d_skip, g, counters, sidx, skey);
  nl += 5;
  // level 0
  const uint64_t T0 = (g.max_entries + MSM_K0 - 1) / MSM_K0;
  uint32_t* kk[2] = {ws.pk0.template as<uint32_t>(), ws.pk1.template as<uint32_t>()};
  XYZZ<F>* pp[2] = {ws.pp0.template as<XYZZ<F>>(), ws.pp1.template as<XYZZ<F>>()};
  if (ev_acc0) cudaEventRecord(ev_acc0, st);
  msm_accum_l0<F, MSM_K0><<<(unsigned)((T0 + 127) / 128), 128, 0, st>>>(d_bases, sidx, skey, offsets + g.nkeys, T0, buckets, kk[0], pp[0]);
  if (ev_acc1) cudaEventRecord(ev_acc1, st);
  nl += 1;
  // levels >= 1: ping-pong between the two partial buffers, then one fused tail
  uint64_t S = 2 * T0;
  int cur = 0;
  while (S > (uint64_t)MSM_TAIL_S) {
    const uint64_t T = (S + MSM_KF - 1) / MSM_KF;
    msm_accum_ln<F, MSM_KF><<<(unsigned)((T + 127) / 128), 128, 0, st>>>(kk[cur], pp[cur], S, T, buckets, kk[cur ^ 1], pp[cur ^ 1]);
    nl += 1;
    S = 2 * T;
    cur ^= 1;
  }
  msm_accum_tail<F, MSM_KF><<<1, 256, 0, st>>>(kk[cur], pp[cur], kk[cur ^ 1], pp[cur ^ 1], S, buckets);
  nl += 1;
  // bucket reduction
  const uint32_t L = msm_seg_len(g);
  uint32_t per_win = g.B / L;
  const uint32_t nsegs = (uint32_t)g.ne * per_win;
  XYZZ<F>* a = ws.seg0.template as<XYZZ<F>>();
  XYZZ<F>* b = ws.seg1.template as<XYZZ<F>>();
  msm_bucket_reduce<F><<<(nsegs + 127) / 128, 128, 0, st>>>(buckets, g.B, L, nsegs, a);
  nl += 1;
  while (per_win > 1) {
    const uint32_t R = per_win < (uint32_t)MSM_GRP ? per_win : (uint32_t)MSM_GRP;   // per_win is a power of two
    const uint32_t ng = (uint32_t)g.ne * (per_win / R);
    msm_sum_groups<F><<<(ng + 127) / 128, 128, 0, st>>>(a, R, ng, b);
    nl += 1;
    XYZZ<F>* t = a; a = b; b = t;
    per_win /= R;
  }
  if (ctr) ctr->launches += nl;
  e = cudaMemcpyAsync(ws.h_winsums, a, (size_t)g.ne * sizeof(XYZZ<F>), cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

// Host Horner over the effective-window sums (stream must be synchronised): sum_e 2^(c e) S_e.
template <class F>
XYZZ<F> msm_finish(const MsmWorkspace<F>& ws, const MsmGeom& g) {
  XYZZ<F> acc = XYZZ<F>::inf();
  if (g.n == 0) return acc;
  for (int w = g.ne - 1; w >= 0; w--) {
    for (int k = 0; k < g.c; k++) acc.dbl_inplace();
    acc.add(ws.h_winsums[w]);
  }
  return acc;
}


// Make a query resident: identity mask of copy 0 and (copies > 1) the precomputed multiples 2^(shift*j) * P.
template <class F>
cudaError_t msm_prepare_query(cudaStream_t st, Affine<F>* d_bases, uint32_t cnt, int copies, int shift, uint8_t* d_mask) {
  if (!cnt) return cudaSuccess;
  msm_inf_mask<F><<<(cnt + 255) / 256, 256, 0, st>>>(d_bases, cnt, d_mask);
  if (copies > 1) msm_precompute<F><<<(cnt + 127) / 128, 128, 0, st>>>(d_bases, cnt, copies, shift, d_bases);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// fixed-base batch multiplication (BatchMulPreprocessing::batch_mul, generator.rs:129-183)
// ------------------------------------------------------------------------------------------------
static constexpr int FB_WINDOWS = 32;  // 8-bit windows over a 256-bit scalar
template <class F>
__global__ void fb_table_kernel(Affine<F> g, XYZZ<F>* table /* [32][255] */) {
  const int w = threadIdx.x;
  if (w >= FB_WINDOWS) return;
  XYZZ<F> base = XYZZ<F>::from_affine(g);
  for (int i = 0; i < 8 * w; i++) base.dbl_inplace();
  XYZZ<F> acc = base;
  for (int d = 1; d <= 255; d++) {
    table[w * 255 + d - 1] = acc;
    acc.add(base);
  }
}
template <class F, class FrF>
__global__ void __launch_bounds__(128) fb_mul_kernel(const XYZZ<F>* __restrict__ table, const FrF* __restrict__ scalars,
                                                     uint32_t n, Affine<F>* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  FrF s;
  {
    const uint4* p = reinterpret_cast<const uint4*>(scalars + i);
    uint4 lo = __ldg(p), hi = __ldg(p + 1);
    s.v[0] = lo.x; s.v[1] = lo.y; s.v[2] = lo.z; s.v[3] = lo.w;
    s.v[4] = hi.x; s.v[5] = hi.y; s.v[6] = hi.z; s.v[7] = hi.w;
  }
  s = FrF::from_mont(s);
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int w = 0; w < FB_WINDOWS; w++) {
    const uint32_t d = (s.v[w >> 2] >> (8 * (w & 3))) & 0xff;
    if (d) acc.add(table[w * 255 + d - 1]);
  }
  out[i] = acc.to_affine();
}
// d_table: FB_WINDOWS * 255 XYZZ points of scratch
template <class F, class FrF>
cudaError_t fb_batch_mul(cudaStream_t st, const Affine<F>& gen, const FrF* d_scalars, uint64_t cnt, Affine<F>* d_out,
                         XYZZ<F>* d_table) {
  fb_table_kernel<F><<<1, 32, 0, st>>>(gen, d_table);
  if (cnt) fb_mul_kernel<F, FrF><<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(d_table, d_scalars, (uint32_t)cnt, d_out);
  return cudaGetLastError();
}

// Explicit-instantiation lists: kernels are compiled in their own translation units (k_msm_*.cu), the engine TU only
// declares them `extern template` (keeps ptxas work parallel across make jobs).
#define G16_MSM_TEMPLATES(X, F, FrF)                                                                                     \
  X cudaError_t msm_enqueue<F, FrF>(cudaStream_t, MsmWorkspace<F>&, const MsmGeom&, const Affine<F>*, const uint8_t*,    \
                                    const uint32_t*, bool, MsmCounters*, cudaEvent_t, cudaEvent_t);                      \
  X cudaError_t msm_prepare_query<F>(cudaStream_t, Affine<F>*, uint32_t, int, int, uint8_t*);                            \
  X cudaError_t fb_batch_mul<F, FrF>(cudaStream_t, const Affine<F>&, const FrF*, uint64_t, Affine<F>*, XYZZ<F>*);

}  // namespace g16
