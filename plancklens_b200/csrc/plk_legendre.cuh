// Legendre stage of the spin-weighted SHT for sm_100a: seed (ramp-up) kernel, synthesis and analysis kernels.
//
// Work decomposition: a thread owns NR iso-latitude ring PAIRS (north ring + southern twin share the
// recurrence through  slam_lm(pi-theta) = (-1)^{l+m} (-s)lam_lm(theta) ), a block owns one m and a tile of
// NCW*32*NR pairs, and walks l in chunks.  Per-l data (pre-scaled a_lm and the recurrence coefficients U,V)
// are staged global->shared with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) by a producer
// warp through a NSTAGE ring; the compute warps read them as broadcast LDS.128.
//
// FP64 pipe cost per (l, ring pair): spin 0: 1 DMUL + 3 DFMA ; spin s: 12 DFMA (2 coefficient, 2 recurrence,
// 8 accumulate) -- the figures SURVEY.md section 8(d) uses for the roofline.
#pragma once
#include <cuda_runtime.h>

#include "plk_common.h"

namespace plk {

#ifndef PLK_PRODUCER_WARP
#define PLK_PRODUCER_WARP 0   // 1: a fifth warp only issues the TMA copies (round-1 layout); 0: lane 0 of warp 0 issues them
#endif                        //    between chunks -- the block shrinks to 128 threads and a third block fits each SM
#ifndef PLK_CHUNK
#define PLK_CHUNK 128
#endif
#ifndef PLK_SYN_MINB
#define PLK_SYN_MINB (PLK_PRODUCER_WARP ? 1 : 3)   // 128-thread blocks: three per SM if <= 170 registers
#endif
#ifndef PLK_ANA_MINB
#define PLK_ANA_MINB (PLK_PRODUCER_WARP ? 1 : 3)
#endif
#ifndef PLK_ANA_PIPE
#define PLK_ANA_PIPE 1     // software-pipelined butterfly in the analysis kernel (0: plain loop)
#endif
#ifndef PLK_ANA_MINB_S2
#define PLK_ANA_MINB_S2 (PLK_PRODUCER_WARP ? 2 : 3)  // resident blocks asked for the spin-s NR = 2 analysis kernel
#endif
constexpr int kChunk = PLK_CHUNK;      // l values per TMA stage (even)
constexpr int kStages = 4;
constexpr int kNCW = 4;          // compute warps per block
constexpr int kLegThreads = (kNCW + PLK_PRODUCER_WARP) * 32;
constexpr int kSeedThrExp = -60;   // default start threshold: accumulation starts once |p_l| >= 2^-60, libsharp's own
                                   // sharp_ftol.  Measured at nside = lmax = 2048 against 2^-200: rel. L2 3e-17, max-abs 6e-16
                                   // of the largest value (2^-120 and 2^-90: bit-identical to 2^-200) while the walked volume
                                   // drops from 0.739 to 0.711 (scripts/time_threshold.py).  Per plan through
                                   // plk_plan_set_seed_threshold (the parity tests vary it)

struct DevGeom {
  int nside, npair, nring;
  long long npix;
  const double *cth;
  const double *sh_hi, *sh_lo, *ch_hi, *ch_lo;
};

struct DevSpin {
  int spin, lmax, mmax;
  int thr_exp;             // start threshold exponent (kSeedThrExp unless the plan overrides it)
  const double2 *UV;       // [alm_idx]
  const double *alpha;     // [alm_idx]
  const double *k_hi, *k_lo;
  const int *k_e, *pc, *ps;
  const signed char *sg_p, *sg_m;
  // seeds, [m * npair + ip]
  int *ks;                 // first accumulated offset k = l - l0(m) (even); K(m) or more = skip
  double *s0, *s1;         // p^+_{l-1}, p^+_l at l = l0 + ks
  double *s2, *s3;         // p^-  (spin > 0)
  int *mtop;               // [npair] highest m with a non-skipped seed on this pair (-1: none)
};

// ---------------------------------------------------------------------------------------------- PTX helpers
PLK_D uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
PLK_D void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
PLK_D void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
PLK_D void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
PLK_D void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
PLK_D void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  const uint32_t a = smem_u32(bar);
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes or ~hint_ns pass
PLK_D bool mbar_try_hint(uint64_t *bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
// 1-D bulk async copy global -> shared (TMA engine), completion counted in bytes on an mbarrier.
PLK_D void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
PLK_D void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---------------------------------------------------------------------------------------------- seeds
// One thread per (m, ring pair): evaluates the closed-form d^{l0} seed in extended-range double-double, runs the
// normalised recurrence with a power-of-two scale until the true value is >= 2^kSeedThrExp, and stores the
// (even) start offset together with the two (four) plain-double start values.
template <bool SPIN>
PLK_HD void seed_one(const DevGeom &g, const DevSpin &t, int m, int ip, int &ks_out, double &o0, double &o1, double &o2,
                     double &o3) {
  const int s = t.spin, lmax = t.lmax;
  const int l0 = m > s ? m : s;
  const int K = lmax - l0 + 1;
  ks_out = K > 0 ? K : 0; o0 = o1 = o2 = o3 = 0.0;
  if (K <= 0) return;
  xdd ch = xdd_norm(g.ch_hi[ip], g.ch_lo[ip], 0);
  xdd sh = xdd_norm(g.sh_hi[ip], g.sh_lo[ip], 0);
  xdd kap; kap.hi = t.k_hi[m]; kap.lo = t.k_lo[m]; kap.e = t.k_e[m];
  const int pc = t.pc[m], ps = t.ps[m];
  xdd cpc = xdd_pow(ch, pc), sps = xdd_pow(sh, ps);
  xdd sp = xdd_mul(kap, xdd_mul(cpc, sps));
  double ap_m = 0.0, ap_c = (sp.hi + sp.lo) * (double)t.sg_p[m];
  int ep = sp.e;
  double am_m = 0.0, am_c = 0.0;
  int em = 0;
  if (SPIN) {
    xdd cps = xdd_pow(ch, ps), spc = xdd_pow(sh, pc);
    xdd sm = xdd_mul(kap, xdd_mul(cps, spc));
    am_c = (sm.hi + sm.lo) * (double)t.sg_m[m];
    em = sm.e;
  }
  const double x = g.cth[ip];
  const double2 *UV = t.UV + alm_idx(lmax, 0, m);
  const double big = ldexp(1.0, 128), small = ldexp(1.0, -128);
  int k = 0;
  for (;;) {
    if ((k & 1) == 0) {
      // true magnitudes: |a_c| * 2^e ; a_c in [2^-129, 2^128] (or 0)
      bool okp = (ap_c != 0.0) && (ilogb(ap_c) + ep >= t.thr_exp);
      bool okm = SPIN && (am_c != 0.0) && (ilogb(am_c) + em >= t.thr_exp);
      if (okp || okm) break;
    }
    if (k >= K - 1) { ks_out = K; return; }   // never reaches the threshold inside the band: pair skipped for this m
    const double2 uv = UV[l0 + k];
    {
      double tp = fma(x, uv.x, uv.y);
      double n = fma(tp, ap_c, -ap_m);
      ap_m = ap_c; ap_c = n;
      if (fabs(ap_c) > big) { ap_c *= small; ap_m *= small; ep += 128; }
    }
    if (SPIN) {
      double tm = fma(x, uv.x, -uv.y);
      double n = fma(tm, am_c, -am_m);
      am_m = am_c; am_c = n;
      if (fabs(am_c) > big) { am_c *= small; am_m *= small; em += 128; }
    }
    ++k;
  }
  ks_out = k;
  // ldexp of a denormal-range result flushes gracefully; both sequences are > 2^-400 here by construction
  o0 = ldexp(ap_m, ep); o1 = ldexp(ap_c, ep);
  if (SPIN) { o2 = ldexp(am_m, em); o3 = ldexp(am_c, em); }
}

template <bool SPIN>
__global__ void seed_kernel(DevGeom g, DevSpin t) {
  const int ip = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (ip >= g.npair || m > t.mmax) return;
  int ks; double a, b, c, d;
  seed_one<SPIN>(g, t, m, ip, ks, a, b, c, d);
  const size_t o = (size_t)m * g.npair + ip;
  t.ks[o] = ks; t.s0[o] = a; t.s1[o] = b;
  if (SPIN) { t.s2[o] = c; t.s3[o] = d; }
}

// highest m each ring pair takes part in (the ring-FFT stage folds / unfolds nothing above it)
__global__ void mtop_kernel(DevGeom g, DevSpin t) {
  const int ip = blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= g.npair) return;
  int top = -1;
  for (int m = t.mmax; m >= 0; --m) {
    const int l0 = m > t.spin ? m : t.spin;
    if (t.ks[(size_t)m * g.npair + ip] < t.lmax - l0 + 1) { top = m; break; }
  }
  t.mtop[ip] = top;
}

// ---------------------------------------------------------------------------------------------- shared layout
template <bool SPIN, bool SYNTH>
struct StageBytes {
  // synthesis spin0: a' (16 B) + UV (16 B); spin: H+,H- (32 B) + UV (16 B); analysis: UV only
  static constexpr int rec = SYNTH ? (SPIN ? 32 : 16) : 0;
  static constexpr int per_l = rec + 16;
  static constexpr int stage = per_l * kChunk;
};

// The synthesis loop takes the l of a row two at a time: behind a chunk that ends on an odd count (only the last chunk of a
// row of odd length can: kChunk is even) the producer appends one ZERO record and one zero UV pair, copied by the same
// engine onto the same mbarrier as the data -- the compute warps never write to the stage themselves.
__device__ __align__(16) double g_zero_rec[4] = {0.0, 0.0, 0.0, 0.0};   // never written
template <bool SPIN, bool SYNTH>
PLK_D uint32_t pad_bytes(int n) { return (SYNTH && (n & 1)) ? (uint32_t)(StageBytes<SPIN, SYNTH>::rec + 16) : 0u; }
template <bool SPIN, bool SYNTH>
PLK_D void pad_odd_row(unsigned char *dst, int n, uint64_t *bar) {
  using SB = StageBytes<SPIN, SYNTH>;
  if (!SYNTH || !(n & 1)) return;
  tma_load_1d(dst + (size_t)n * SB::rec, g_zero_rec, (uint32_t)SB::rec, bar);
  tma_load_1d(dst + SB::rec * kChunk + (size_t)n * 16, g_zero_rec, 16u, bar);
}

// Producer warp body: streams chunks [c0, nchunk) of row m into the stage ring.  One lane works; while the
// ring is full it sleeps (nanosleep back-off) instead of burning issue slots the FP64 warps need.
template <bool SPIN, bool SYNTH>
PLK_D void producer_loop(unsigned char *stage_base, uint64_t *full, uint64_t *empty, const void *rec_row,
                         const double2 *uv_row, int K, int c0, int nchunk) {
  using SB = StageBytes<SPIN, SYNTH>;
  if ((threadIdx.x & 31) != 0) return;
  for (int c = c0; c < nchunk; ++c) {
    const int it = c - c0;
    const int st = it % kStages;
    if (it >= kStages) {
      const uint32_t par = ((it / kStages) - 1) & 1;
      while (!mbar_try_hint(&empty[st], par, 100000u)) {}
    }
    const int k0 = c * kChunk;
    const int n = min(kChunk, K - k0);
    unsigned char *dst = stage_base + (size_t)st * SB::stage;
    const uint32_t bytes = (uint32_t)n * SB::per_l;
    mbar_expect_tx(&full[st], bytes + pad_bytes<SPIN, SYNTH>(n));
    if (SYNTH) tma_load_1d(dst, (const unsigned char *)rec_row + (size_t)k0 * SB::rec, (uint32_t)n * SB::rec, &full[st]);
    tma_load_1d(dst + SB::rec * kChunk, uv_row + k0, (uint32_t)n * 16, &full[st]);
    pad_odd_row<SPIN, SYNTH>(dst, n, &full[st]);
  }
}

// One chunk of row m into stage (c - c0) % kStages; the caller has made sure the stage is free.
template <bool SPIN, bool SYNTH>
PLK_D void issue_chunk(unsigned char *stage_base, uint64_t *full, const void *rec_row, const double2 *uv_row, int K,
                       int c0, int c) {
  using SB = StageBytes<SPIN, SYNTH>;
  const int st = (c - c0) % kStages;
  const int k0 = c * kChunk;
  const int n = min(kChunk, K - k0);
  unsigned char *dst = stage_base + (size_t)st * SB::stage;
  mbar_expect_tx(&full[st], (uint32_t)n * SB::per_l + pad_bytes<SPIN, SYNTH>(n));
  if (SYNTH) tma_load_1d(dst, (const unsigned char *)rec_row + (size_t)k0 * SB::rec, (uint32_t)n * SB::rec, &full[st]);
  tma_load_1d(dst + SB::rec * kChunk, uv_row + k0, (uint32_t)n * 16, &full[st]);
  pad_odd_row<SPIN, SYNTH>(dst, n, &full[st]);
}
// Producer duties folded into compute warp 0 (PLK_PRODUCER_WARP == 0).  Called by every thread of warp 0 after it
// has released chunk c: refills the stage of chunk c - 1 -- which the other warps have almost always left by now,
// so the wait is free -- with chunk c - 1 + kStages (prefetch distance kStages - 1 chunks).
template <bool SPIN, bool SYNTH>
PLK_D void refill_behind(unsigned char *stage_base, uint64_t *full, uint64_t *empty, const void *rec_row,
                         const double2 *uv_row, int K, int c0, int nchunk, int c, int lane) {
  const int cb = c - 1;                       // chunk whose stage gets refilled
  const int cn = cb + kStages;                // chunk to load
  if (cb < c0 || cn >= nchunk) return;
  if (lane == 0) {
    const int it = cb - c0;
    mbar_wait(&empty[it % kStages], (it / kStages) & 1);
    issue_chunk<SPIN, SYNTH>(stage_base, full, rec_row, uv_row, K, c0, cn);
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------- synthesis
// rec rows: spin 0: double2 a'_l = alpha_l * fl_l * a_lm ; spin s: double4 {H+ re, H+ im, H- re, H- im} with
//   H+ = -1/2 alpha (G + iC),  H- = -1/2 (-1)^s alpha (G - iC)        (prepared by prep_alm_kernel)
// output phase arrays X1 (, X2): [ring][pitch] complex, map(phi) = X_0 + 2 Re sum_{m>0} X_m e^{i m phi}
// GRAD (spin s only): the curl input is identically zero (the gradient legs of the quadratic estimators,
//   qest.py:566-595), so H- = (-1)^s H+ and the four sums collapse to even / odd partial sums of H+ p+ and H+ p-:
//   8 instead of 12 DFMA per (l, ring pair).
template <bool SPIN, int NR, bool GRAD = false>
__global__ void __launch_bounds__(kLegThreads, PLK_SYN_MINB)
legendre_synth_kernel(DevGeom g, DevSpin t, const void *__restrict__ rec, cplx *__restrict__ X1, cplx *__restrict__ X2,
                      int pitch, const int *__restrict__ morder, DistX dx) {
  using SB = StageBytes<SPIN, true>;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);
  uint64_t *empty = full + kStages;
  int *s_kmin = reinterpret_cast<int *>(empty + kStages);
  unsigned char *stage_base = smem + 128;
  // start values of every ring pair of the block, thread-private slots [j][component][thread]: the in-loop
  // injection then costs one LDS instead of a divergent global load that stalls the whole warp
  double *sseed = reinterpret_cast<double *>(stage_base + (size_t)kStages * SB::stage);
  constexpr int NSD = SPIN ? 4 : 2;

  const int m = morder[blockIdx.y];
  const int s = t.spin, lmax = t.lmax;
  const int l0 = m > s ? m : s;
  const int K = lmax - l0 + 1;
  if (K <= 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_pairs = kNCW * 32 * NR;
  const int pair0 = blockIdx.x * tile_pairs + warp * 32 * NR + lane;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kNCW); }
    mbar_fence_init();
    *s_kmin = 1 << 30;
  }
  __syncthreads();

  // per-thread ring-pair state
  int ks[NR];
  double x[NR];
  int kw_min = 1 << 30, kw_max = -1;
  double *sd = sseed + threadIdx.x;
  if (warp < kNCW) {
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      const int ip = pair0 + 32 * j;
      if (ip < g.npair) {
        const size_t so = (size_t)m * g.npair + ip;
        ks[j] = t.ks[so];
        x[j] = g.cth[ip];
        if (ks[j] < K) {
          sd[(j * NSD + 0) * (kNCW * 32)] = t.s0[so];
          sd[(j * NSD + 1) * (kNCW * 32)] = t.s1[so];
          if (SPIN) { sd[(j * NSD + 2) * (kNCW * 32)] = t.s2[so]; sd[(j * NSD + 3) * (kNCW * 32)] = t.s3[so]; }
        }
      } else { ks[j] = 1 << 30; x[j] = 0.0; }
      if (ks[j] < K) { kw_min = min(kw_min, ks[j]); kw_max = max(kw_max, ks[j]); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      kw_min = min(kw_min, __shfl_xor_sync(0xffffffffu, kw_min, o));
      kw_max = max(kw_max, __shfl_xor_sync(0xffffffffu, kw_max, o));
    }
    if (lane == 0) atomicMin(s_kmin, kw_min);
  }
  __syncthreads();
  const int kb_min = *s_kmin;
  const size_t row = (size_t)alm_idx(lmax, l0, m);
  const int nchunk = (K + kChunk - 1) / kChunk;
  const int c0 = kb_min >= K ? nchunk : kb_min / kChunk;

  const unsigned char *rec_row = (const unsigned char *)rec + row * SB::rec;
  if (PLK_PRODUCER_WARP) {
    if (warp == kNCW) {
      producer_loop<SPIN, true>(stage_base, full, empty, rec_row, t.UV + row, K, c0, nchunk);
      return;
    }
  } else if (threadIdx.x == 0) {
    for (int c = c0; c < min(nchunk, c0 + kStages); ++c) issue_chunk<SPIN, true>(stage_base, full, rec_row, t.UV + row, K, c0, c);
  }

  // accumulators
  double pc_[NR], pm_[NR], qc_[NR], qm_[NR];
  double a0r[NR], a0i[NR], a1r[NR], a1i[NR];   // spin0: even / odd ; spin: north A+ (P), north A- (M)
  double b0r[NR], b0i[NR], b1r[NR], b1i[NR];   // spin: south A+ (uses p-), south A- (uses p+)
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    pc_[j] = pm_[j] = qc_[j] = qm_[j] = 0.0;
    a0r[j] = a0i[j] = a1r[j] = a1i[j] = b0r[j] = b0i[j] = b1r[j] = b1i[j] = 0.0;
  }

  for (int c = c0; c < nchunk; ++c) {
    const int it = c - c0;
    const int st = it % kStages;
    mbar_wait(&full[st], (it / kStages) & 1);
    const int k0 = c * kChunk;
    const int kend = min(k0 + kChunk, K);
    unsigned char *sb = stage_base + (size_t)st * SB::stage;
    // (odd row length: the (k, k+1) loop below reads one record past the row -- a zero record the producer appended)
    if (kend > kw_min) {
      const double2 *uvs = reinterpret_cast<const double2 *>(sb + SB::rec * kChunk);
      for (int k = max(k0, kw_min); k < kend; k += 2) {
        if (k <= kw_max) {
#pragma unroll
          for (int j = 0; j < NR; ++j)
            if (k == ks[j]) {
              pm_[j] = sd[(j * NSD + 0) * (kNCW * 32)]; pc_[j] = sd[(j * NSD + 1) * (kNCW * 32)];
              if (SPIN) { qm_[j] = sd[(j * NSD + 2) * (kNCW * 32)]; qc_[j] = sd[(j * NSD + 3) * (kNCW * 32)]; }
            }
        }
        const int kk = k - k0;
        if (!SPIN) {
          const double2 *as = reinterpret_cast<const double2 *>(sb);
          const double2 e = as[kk], o = as[kk + 1];
          const double ue = uvs[kk].x, uo = uvs[kk + 1].x;
#pragma unroll
          for (int j = 0; j < NR; ++j) {
            a0r[j] = fma(e.x, pc_[j], a0r[j]); a0i[j] = fma(e.y, pc_[j], a0i[j]);
            const double n1 = fma(x[j] * ue, pc_[j], -pm_[j]);
            a1r[j] = fma(o.x, n1, a1r[j]); a1i[j] = fma(o.y, n1, a1i[j]);
            const double n2 = fma(x[j] * uo, n1, -pc_[j]);
            pm_[j] = n1; pc_[j] = n2;
          }
        } else {
          const double4 *hs = reinterpret_cast<const double4 *>(sb);
          const double4 he = hs[kk], ho = hs[kk + 1];
          const double2 ue = uvs[kk], uo = uvs[kk + 1];
#pragma unroll
          for (int j = 0; j < NR; ++j) {
            if (GRAD) {
              // a0 = E+ (even l, H+ p+), b1 = O+ (odd l, H+ p+), a1 = E- (even, H+ p-), b0 = O- (odd, H+ p-)
              a0r[j] = fma(he.x, pc_[j], a0r[j]); a0i[j] = fma(he.y, pc_[j], a0i[j]);
              a1r[j] = fma(he.x, qc_[j], a1r[j]); a1i[j] = fma(he.y, qc_[j], a1i[j]);
              const double np = fma(fma(x[j], ue.x, ue.y), pc_[j], -pm_[j]);
              const double nq = fma(fma(x[j], ue.x, -ue.y), qc_[j], -qm_[j]);
              b1r[j] = fma(ho.x, np, b1r[j]); b1i[j] = fma(ho.y, np, b1i[j]);
              b0r[j] = fma(ho.x, nq, b0r[j]); b0i[j] = fma(ho.y, nq, b0i[j]);
              const double np2 = fma(fma(x[j], uo.x, uo.y), np, -pc_[j]);
              const double nq2 = fma(fma(x[j], uo.x, -uo.y), nq, -qc_[j]);
              pm_[j] = np; pc_[j] = np2; qm_[j] = nq; qc_[j] = nq2;
              continue;
            }
            // even offset: sigma = +1
            a0r[j] = fma(he.x, pc_[j], a0r[j]); a0i[j] = fma(he.y, pc_[j], a0i[j]);   // north A+ += H+ p+
            b1r[j] = fma(he.z, pc_[j], b1r[j]); b1i[j] = fma(he.w, pc_[j], b1i[j]);   // south A- += H- p+
            a1r[j] = fma(he.z, qc_[j], a1r[j]); a1i[j] = fma(he.w, qc_[j], a1i[j]);   // north A- += H- p-
            b0r[j] = fma(he.x, qc_[j], b0r[j]); b0i[j] = fma(he.y, qc_[j], b0i[j]);   // south A+ += H+ p-
            const double np = fma(fma(x[j], ue.x, ue.y), pc_[j], -pm_[j]);
            const double nq = fma(fma(x[j], ue.x, -ue.y), qc_[j], -qm_[j]);
            // odd offset: sigma = -1 on the southern sums
            a0r[j] = fma(ho.x, np, a0r[j]); a0i[j] = fma(ho.y, np, a0i[j]);
            b1r[j] = fma(-ho.z, np, b1r[j]); b1i[j] = fma(-ho.w, np, b1i[j]);
            a1r[j] = fma(ho.z, nq, a1r[j]); a1i[j] = fma(ho.w, nq, a1i[j]);
            b0r[j] = fma(-ho.x, nq, b0r[j]); b0i[j] = fma(-ho.y, nq, b0i[j]);
            const double np2 = fma(fma(x[j], uo.x, uo.y), np, -pc_[j]);
            const double nq2 = fma(fma(x[j], uo.x, -uo.y), nq, -qc_[j]);
            pm_[j] = np; pc_[j] = np2; qm_[j] = nq; qc_[j] = nq2;
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
    if (!PLK_PRODUCER_WARP && warp == 0)
      refill_behind<SPIN, true>(stage_base, full, empty, rec_row, t.UV + row, K, c0, nchunk, c, lane);
  }

  // sigma at even offset is (-1)^{l0+m}
  const double sg0 = ((l0 + m) & 1) ? -1.0 : 1.0;
  if (GRAD) {
    // north A+ = E+ + O+, south A- = sg (E+ - O+), north A- = sg (E- + O-), south A+ = E- - O-,  sg = (-1)^s
    const double sgs = (s & 1) ? -1.0 : 1.0;
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      const double epr = a0r[j], epi = a0i[j], opr = b1r[j], opi = b1i[j];
      const double emr = a1r[j], emi = a1i[j], omr = b0r[j], omi = b0i[j];
      a0r[j] = epr + opr; a0i[j] = epi + opi;
      b1r[j] = sgs * (epr - opr); b1i[j] = sgs * (epi - opi);
      a1r[j] = sgs * (emr + omr); a1i[j] = sgs * (emi + omi);
      b0r[j] = emr - omr; b0i[j] = emi - omi;
    }
  }
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    const int ip = pair0 + 32 * j;
    if (ip >= g.npair) continue;
    const int rn = ip, rs = g.nring - 1 - ip;
    if (dx.nranks > 1) {
      // m-partitioned transform: this rank holds only its m columns; the values go straight into the exchange
      // buffer of the rank that owns the ring pair (peer-memory stores over NVLink -- the all-to-all is fused into
      // the kernel).  The exchange buffer is TRANSPOSED, T[m][ring]: the 32 ring pairs of a warp then form one
      // contiguous 512-byte run per store instruction instead of 32 isolated 16-byte NVLink writes
      // (measured: 2.8x slower Legendre stage with the [ring][m] layout); the owner transposes locally.
      int q = 0;
      while (q + 1 < dx.nranks && ip >= dx.pair_lo[q + 1]) ++q;
      cplx *T1 = dx.x1[q] + (size_t)m * dx.tpitch, *T2 = dx.x2[q] + (size_t)m * dx.tpitch;
      if (!SPIN) {
        T1[rn] = mk(a0r[j] + a1r[j], a0i[j] + a1i[j]);
        if (rs != rn) T1[rs] = mk(sg0 * (a0r[j] - a1r[j]), sg0 * (a0i[j] - a1i[j]));
      } else {
        T1[rn] = mk(a0r[j] + a1r[j], a0i[j] + a1i[j]);
        T2[rn] = mk(a0i[j] - a1i[j], -(a0r[j] - a1r[j]));
        if (rs != rn) {
          T1[rs] = mk(sg0 * (b0r[j] + b1r[j]), sg0 * (b0i[j] + b1i[j]));
          T2[rs] = mk(sg0 * (b0i[j] - b1i[j]), -sg0 * (b0r[j] - b1r[j]));
        }
      }
      continue;
    }
    if (!SPIN) {
      X1[(size_t)rn * pitch + m] = mk(a0r[j] + a1r[j], a0i[j] + a1i[j]);
      if (rs != rn) X1[(size_t)rs * pitch + m] = mk(sg0 * (a0r[j] - a1r[j]), sg0 * (a0i[j] - a1i[j]));
    } else {
      // X1 = P + M ; X2 = -i (P - M)
      X1[(size_t)rn * pitch + m] = mk(a0r[j] + a1r[j], a0i[j] + a1i[j]);
      X2[(size_t)rn * pitch + m] = mk(a0i[j] - a1i[j], -(a0r[j] - a1r[j]));
      if (rs != rn) {
        X1[(size_t)rs * pitch + m] = mk(sg0 * (b0r[j] + b1r[j]), sg0 * (b0i[j] + b1i[j]));
        X2[(size_t)rs * pitch + m] = mk(sg0 * (b0i[j] - b1i[j]), -sg0 * (b0r[j] - b1r[j]));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- analysis
// Input phase arrays X1 (, X2): X_m(ring) = sum_j map_j e^{-i m phi_j} (quadrature weight already applied).
// Output partial sums part[tile][alm_idx(l,m)][NV] (NV = 2 doubles spin 0, 4 doubles spin s):
//   spin 0:  S_l   = sum_pairs p_l (X_n + sigma_l X_s)
//   spin s:  S+_l  = sum_pairs [p+ F+_n + sigma_l p- F+_s],  S-_l = sum_pairs [p- F-_n + sigma_l p+ F-_s],
//            F+- = X1 +- i X2.   alpha and the G/C recombination are applied by finish_alm_kernel, which also
//            reduces over tiles (deterministic order).
// Cross-lane reduction: 16 accumulators per lane, index = vs * NB + b (component slot high, l offset low), then
// a butterfly reduce-scatter (8+4+2+1+1 shuffles).  The slot -> component map is permuted per lane
// (component = vs ^ lane bits) by permuting the ring constants once at load time, so the upper butterfly levels
// need no selects: every lane keeps its low registers and receives the partner's high ones.
template <int NSEL>   // number of leading select-free levels (2 for spin s, 1 for spin 0)
PLK_D void butterfly16(double (&v)[16], int lane) {
  // levels bit = 16, 8, 4, 2 halve the live registers; the last level (bit 1) is a plain sum
#pragma unroll
  for (int lev = 0; lev < 4; ++lev) {
    const int h = 8 >> lev, bit = 16 >> lev;
    if (lev < NSEL) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < h) v[j] += __shfl_xor_sync(0xffffffffu, v[j + h], bit);
    } else {
      const bool up = (lane & bit) != 0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < h) {
          const double send = up ? v[j] : v[j + h];
          const double keep = up ? v[j + h] : v[j];
          v[j] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

template <bool SPIN, int NR>
// NR = 4: two resident blocks (255 registers, no spills) beat three with NR = 2 -- half the butterfly shuffles per FMA:
// spin s 8.21 -> 7.67 ms, spin 0 2.86 -> 2.82 ms at nside = lmax = 2048 (variants a / NR 4, round 2)
__global__ void __launch_bounds__(kLegThreads, NR == 4 ? 2 : ((SPIN && NR == 2) ? PLK_ANA_MINB_S2 : PLK_ANA_MINB))
legendre_anal_kernel(DevGeom g, DevSpin t, const cplx *__restrict__ X1, const cplx *__restrict__ X2, int pitch,
                     double *__restrict__ part, long long part_stride /* doubles per tile */,
                     const int *__restrict__ morder, int dbg) {
  using SB = StageBytes<SPIN, false>;
  constexpr int NV = SPIN ? 4 : 2;
  constexpr int NB = 16 / NV;   // l values per butterfly batch
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);
  uint64_t *empty = full + kStages;
  int *s_kmin = reinterpret_cast<int *>(empty + kStages);
  unsigned char *stage_base = smem + 128;
  double *red = reinterpret_cast<double *>(stage_base + (size_t)kStages * SB::stage);  // [2][kNCW][kChunk*NV]
  double *sseed = red + (size_t)2 * kNCW * kChunk * NV;     // [j][component][thread] start values (see synthesis)
  constexpr int NSD = SPIN ? 4 : 2;
  constexpr bool SEED_SMEM = !SPIN;   // measured on B200: helps spin 0 (4.47 -> 3.98 ms), hurts spin s (10.4 -> 11.6 ms)

  const int m = morder[blockIdx.y];
  const int s = t.spin, lmax = t.lmax;
  const int l0 = m > s ? m : s;
  const int K = lmax - l0 + 1;
  if (K <= 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_pairs = kNCW * 32 * NR;
  const int pair0 = blockIdx.x * tile_pairs + warp * 32 * NR + lane;
  const int hi = SPIN ? ((lane >> 3) & 3) : ((lane >> 4) & 1);   // lane bits consumed by the select-free levels

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kNCW); }
    mbar_fence_init();
    *s_kmin = 1 << 30;
  }
  __syncthreads();

  int ks[NR];
  double x[NR];
  size_t so_[NR];
  double *sd = sseed + threadIdx.x;
  // lane-permuted ring constants.  spin s: fa[vs] multiplies p+, fb[vs] multiplies p-, component v = vs ^ hi of
  //   (S+re, S+im, S-re, S-im):  A = (F+n.re, F+n.im, sF-s.re, sF-s.im), B = (sF+s.re, sF+s.im, F-n.re, F-n.im)
  // spin 0: fa[vs] = (fe.re, fe.im)[vs ^ hi] used at even offsets, fb[vs] = (fo.re, fo.im)[vs ^ hi] at odd ones
  double fa[NR][NV], fb[NR][NV];
  int kw_min = 1 << 30, kw_max = -1;
  const double sg0 = ((l0 + m) & 1) ? -1.0 : 1.0;
  if (warp < kNCW) {
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      const int ip = pair0 + 32 * j;
      double A[4] = {0.0, 0.0, 0.0, 0.0}, B[4] = {0.0, 0.0, 0.0, 0.0};
      if (ip < g.npair) {
        const size_t so = (size_t)m * g.npair + ip;
        ks[j] = t.ks[so];
        x[j] = g.cth[ip];
        const int rn = ip, rs = g.nring - 1 - ip;
        so_[j] = so;
        if (ks[j] < K) {
          if (SEED_SMEM) {
            sd[(j * NSD + 0) * (kNCW * 32)] = t.s0[so];
            sd[(j * NSD + 1) * (kNCW * 32)] = t.s1[so];
          }
          const cplx n1 = X1[(size_t)rn * pitch + m];
          const cplx s1 = (rs != rn) ? X1[(size_t)rs * pitch + m] : mk(0.0, 0.0);
          if (!SPIN) {
            A[0] = n1.x + sg0 * s1.x; A[1] = n1.y + sg0 * s1.y;     // fe
            B[0] = n1.x - sg0 * s1.x; B[1] = n1.y - sg0 * s1.y;     // fo
          } else {
            const cplx n2 = X2[(size_t)rn * pitch + m];
            const cplx s2 = (rs != rn) ? X2[(size_t)rs * pitch + m] : mk(0.0, 0.0);
            // F+ = X1 + i X2 ; F- = X1 - i X2
            A[0] = n1.x - n2.y; A[1] = n1.y + n2.x;                         // F+ north
            B[2] = n1.x + n2.y; B[3] = n1.y - n2.x;                         // F- north
            B[0] = sg0 * (s1.x - s2.y); B[1] = sg0 * (s1.y + s2.x);         // sigma0 F+ south
            A[2] = sg0 * (s1.x + s2.y); A[3] = sg0 * (s1.y - s2.x);         // sigma0 F- south
          }
        }
      } else { ks[j] = 1 << 30; x[j] = 0.0; so_[j] = 0; }
#pragma unroll
      for (int vs = 0; vs < NV; ++vs) {
        // static-index selection of A[vs ^ hi]
        double av = A[vs], bv = B[vs];
#pragma unroll
        for (int h2 = 1; h2 < NV; ++h2)
          if (hi == h2) { av = A[vs ^ h2]; bv = B[vs ^ h2]; }
        fa[j][vs] = av; fb[j][vs] = bv;
      }
      if (ks[j] < K) { kw_min = min(kw_min, ks[j]); kw_max = max(kw_max, ks[j]); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      kw_min = min(kw_min, __shfl_xor_sync(0xffffffffu, kw_min, o));
      kw_max = max(kw_max, __shfl_xor_sync(0xffffffffu, kw_max, o));
    }
    if (lane == 0) atomicMin(s_kmin, kw_min);
  }
  __syncthreads();
  const int kb_min = *s_kmin;
  const size_t row = (size_t)alm_idx(lmax, l0, m);
  const int nchunk = (K + kChunk - 1) / kChunk;
  const int c0 = kb_min >= K ? nchunk : kb_min / kChunk;
  double *prow = part + (size_t)blockIdx.x * part_stride + row * NV;

  if (PLK_PRODUCER_WARP) {
    if (warp == kNCW) {
      producer_loop<SPIN, false>(stage_base, full, empty, nullptr, t.UV + row, K, c0, nchunk);
      return;
    }
  } else if (threadIdx.x == 0) {
    for (int c = c0; c < min(nchunk, c0 + kStages); ++c) issue_chunk<SPIN, false>(stage_base, full, nullptr, t.UV + row, K, c0, c);
  }
  // chunks before c0 carry no contribution from this tile: write zeros
  for (int i = threadIdx.x; i < min(c0 * kChunk, K) * NV; i += kNCW * 32) prow[i] = 0.0;

  double pc_[NR], pm_[NR], qc_[NR], qm_[NR];
#pragma unroll
  for (int j = 0; j < NR; ++j) pc_[j] = pm_[j] = qc_[j] = qm_[j] = 0.0;
  // value held by this lane after the butterfly: component vout at l offset bout of the batch
  const int bout = (lane >> 1) & (NB - 1);
  const int vout = hi;
  const bool flip = SPIN && (bout & 1) && (vout >= 2);   // odd offsets accumulate -(S-) (see below)

  for (int c = c0; c < nchunk; ++c) {
    const int it = c - c0;
    const int st = it % kStages;
    mbar_wait(&full[st], (it / kStages) & 1);
    const int k0 = c * kChunk;
    const int kend = min(k0 + kChunk, K);
    double *myred = red + ((size_t)(it & 1) * kNCW + warp) * (kChunk * NV);
    const double2 *uvs = reinterpret_cast<const double2 *>(stage_base + (size_t)st * SB::stage);
    if (kend <= kw_min) {
      for (int i = lane; i < kChunk * NV; i += 32) myred[i] = 0.0;
    } else {
// one (even, odd) pair of l offsets K, K + 1 accumulated into slots B, B + 1 of ACC for all NR ring pairs of the thread;
// reads past kend stay inside the stage buffer, those offsets are never written out
#define PLK_ANA_STEP(ACC, K, B)                                                                                   \
  {                                                                                                               \
    const double2 ue = uvs[(K) - k0], uo = uvs[(K) - k0 + 1];                                                     \
    _Pragma("unroll") for (int j = 0; j < NR; ++j) {                                                              \
      if (!SPIN) {                                                                                                \
        _Pragma("unroll") for (int vs = 0; vs < 2; ++vs) ACC[vs * NB + (B)] = fma(pc_[j], fa[j][vs], ACC[vs * NB + (B)]); \
        const double n1 = fma(x[j] * ue.x, pc_[j], -pm_[j]);                                                      \
        _Pragma("unroll") for (int vs = 0; vs < 2; ++vs) ACC[vs * NB + (B) + 1] = fma(n1, fb[j][vs], ACC[vs * NB + (B) + 1]); \
        const double n2 = fma(x[j] * uo.x, n1, -pc_[j]);                                                          \
        pm_[j] = n1; pc_[j] = n2;                                                                                 \
      } else {                                                                                                    \
        /* even offset: component v += p+ A_v + p- B_v */                                                         \
        _Pragma("unroll") for (int vs = 0; vs < 4; ++vs)                                                          \
          ACC[vs * NB + (B)] = fma(pc_[j], fa[j][vs], fma(qc_[j], fb[j][vs], ACC[vs * NB + (B)]));                \
        const double np = fma(fma(x[j], ue.x, ue.y), pc_[j], -pm_[j]);                                            \
        const double nq = fma(fma(x[j], ue.x, -ue.y), qc_[j], -qm_[j]);                                           \
        /* odd offset: S+ += p+ A - p- B and S- += -(p+ A - p- B); the minus on S- is applied after the */      \
        /* butterfly (flip) so that all four slots run the same instruction */                                    \
        _Pragma("unroll") for (int vs = 0; vs < 4; ++vs)                                                          \
          ACC[vs * NB + (B) + 1] = fma(np, fa[j][vs], fma(-nq, fb[j][vs], ACC[vs * NB + (B) + 1]));               \
        const double np2 = fma(fma(x[j], uo.x, uo.y), np, -pc_[j]);                                               \
        const double nq2 = fma(fma(x[j], uo.x, -uo.y), nq, -qc_[j]);                                              \
        pm_[j] = np; pc_[j] = np2; qm_[j] = nq; qc_[j] = nq2;                                                     \
      }                                                                                                           \
    }                                                                                                             \
  }
#define PLK_ANA_BATCH(ACC, KB)                                                                                    \
  {                                                                                                               \
    _Pragma("unroll") for (int i = 0; i < 16; ++i) ACC[i] = 0.0;                                                  \
    _Pragma("unroll") for (int b = 0; b < NB; b += 2) PLK_ANA_STEP(ACC, (KB) + b, b)                              \
  }
#define PLK_ANA_FINISH(ACC, KB)                                                                                   \
  {                                                                                                               \
    butterfly16<SPIN ? 2 : 1>(ACC, lane);                                                                         \
    if ((lane & 1) == 0) myred[((KB) - k0 + bout) * NV + vout] = flip ? -ACC[0] : ACC[0];                         \
  }
      // batches that start beyond the end of the band produce nothing that is read back
      const int kstop = min(k0 + kChunk, (kend + NB - 1) / NB * NB);
      int kb = k0;
      // ramp-up part of the band: some ring pair of the warp still waits for its start offset (seed injection)
      for (; kb < kstop && (PLK_ANA_PIPE == 0 || kb <= kw_max); kb += NB) {
        double acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.0;
        if (kb + NB > kw_min) {
#pragma unroll
          for (int b = 0; b < NB; b += 2) {
            const int k = kb + b;
            if (k <= kw_max && k >= kw_min) {
#pragma unroll
              for (int j = 0; j < NR; ++j)
                if (k == ks[j]) {
                  if (SEED_SMEM) {
                    pm_[j] = sd[(j * NSD + 0) * (kNCW * 32)]; pc_[j] = sd[(j * NSD + 1) * (kNCW * 32)];
                  } else {
                    pm_[j] = t.s0[so_[j]]; pc_[j] = t.s1[so_[j]];
                    if (SPIN) { qm_[j] = t.s2[so_[j]]; qc_[j] = t.s3[so_[j]]; }
                  }
                }
            }
            PLK_ANA_STEP(acc, k, b)
          }
          if (!(dbg & 1)) butterfly16<SPIN ? 2 : 1>(acc, lane);
        }
        if ((lane & 1) == 0) myred[(kb - k0 + bout) * NV + vout] = flip ? -acc[0] : acc[0];
      }
      // steady state: software pipeline -- the cross-lane butterfly of one batch (a chain of five dependent
      // SHFL + DADD levels, ~150 cycles of latency for an in-order warp) is issued next to the independent FMAs of
      // the following batch, both in one basic block so that ptxas interleaves them
      if (PLK_ANA_PIPE != 0 && kb < kstop) {
        double A[16], B[16];
        PLK_ANA_BATCH(A, kb)
        kb += NB;
#pragma unroll 1
        while (kb + NB < kstop) {
          PLK_ANA_BATCH(B, kb)
          PLK_ANA_FINISH(A, kb - NB)
          PLK_ANA_BATCH(A, kb + NB)
          PLK_ANA_FINISH(B, kb)
          kb += 2 * NB;
        }
        if (kb < kstop) {
          PLK_ANA_BATCH(B, kb)
          PLK_ANA_FINISH(A, kb - NB)
          PLK_ANA_FINISH(B, kb)
        } else {
          PLK_ANA_FINISH(A, kb - NB)
        }
      }
#undef PLK_ANA_STEP
#undef PLK_ANA_BATCH
#undef PLK_ANA_FINISH
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
    if (!PLK_PRODUCER_WARP && warp == 0)
      refill_behind<SPIN, false>(stage_base, full, empty, nullptr, t.UV + row, K, c0, nchunk, c, lane);
    if (dbg & 2) continue;
    named_bar_sync(1, kNCW * 32);
    // sum the per-warp slices of this chunk and write the tile partial
    const double *r0 = red + (size_t)(it & 1) * kNCW * (kChunk * NV);
    const int nval = (kend - k0) * NV;
    for (int i = threadIdx.x; i < nval; i += kNCW * 32) {
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < kNCW; ++w) sum += r0[(size_t)w * (kChunk * NV) + i];
      prow[(size_t)k0 * NV + i] = sum;
    }
  }
}

}  // namespace plk
