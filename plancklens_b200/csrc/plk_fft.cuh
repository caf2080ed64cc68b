// Ring FFT stage of the HEALPix SHT for sm_100a: phase array X[ring][m]  <->  RING-ordered pixels.
//
// One thread block per iso-latitude ring PAIR (north ring + southern twin have the same length n = 4q and
// the same phi0).  A real ring of n = 4q samples is handled as two complex length-q DFTs:
//   x_{4t+r} = sum_{k'<q} e^{(r)}_{k'} w_q^{t k'},   e^{(r)}_{k'} = w_n^{r k'} sum_{c<4} i^{rc} d_{k'+qc},
//   packed   y^{(a)}_t = x_{4t+2a} + i x_{4t+2a+1},  a = 0, 1,
// where d_k is the length-n Hermitian spectrum obtained by folding (aliasing) all m <= mmax onto the ring's
// band, including the e^{i m phi0} shift.  The length-q DFT runs entirely in shared memory:
//   q = 2^j        : in-place DIF, output read in bit-reversed order;
//   otherwise      : Bluestein chirp-z with M = nextpow2(2q-1): DIF forward, pointwise multiply by the
//                    precomputed kernel spectrum (stored in DIF output order), DIT inverse -- no bit reversal;
//   q <= kTinyQ    : direct evaluation of the defining sum.
// FFT passes are radix 16 / 8 / 4 in REGISTERS (a thread loads R strided elements, runs log2 R radix-2 levels,
// stores back in place), so a 4096-point transform is 3 shared-memory round trips; the element index is XOR
// swizzled (i ^ ((i >> 4) & 15)) which makes every pass of a radix-16 schedule bank-conflict free.  The 2 or 4
// DFTs of a ring (pair) run as one batch through the same passes.
// Analysis uses the same inverse-DFT machinery on conjugated data (DFT(y) = conj(IDFT(conj y))).
// No cuFFT: 2047 distinct cap-ring lengths per nside would mean thousands of plans and launches per transform.
#pragma once
#include <cuda_runtime.h>

#include "plk_common.h"

namespace plk {

constexpr int kTinyQ = 8;
constexpr int kFftThreads = 256;

// "Pixel program" of the analysis ring kernel: instead of reading one map, pixel p is evaluated on the fly as
//   sum_{k < n} s_k a_k[p] (b_k ? b_k[p] : 1)
// -- the per-pixel products of the quadratic estimators (qest.py:256-257, :276-278) and the N^-1 multiply of the CG
// operators (opfilt_pp.py:272-303, one-map form) fused into the kernel that consumes them.  n == 0: plain map read.
constexpr int kMaxPixTerms = 6;
struct PixProg {
  int n;
  const double *a[kMaxPixTerms], *b[kMaxPixTerms];
  double s[kMaxPixTerms];
};
PLK_HD double pix_eval(const PixProg &q, const double *map, long long p) {
  if (q.n == 0) return map[p];
  double v = 0.0;
  for (int k = 0; k < q.n; ++k) {
    v = fma(q.s[k] * q.a[k][p], q.b[k] ? q.b[k][p] : 1.0, v);
  }
  return v;
}
// four consecutive pixels starting at p (a multiple of 4: every ring starts on a 32-byte boundary)
PLK_HD void pix_eval4(const PixProg &q, long long p, double (&v)[4]) {
  v[0] = v[1] = v[2] = v[3] = 0.0;
  for (int k = 0; k < q.n; ++k) {
    double a[4], b[4] = {1.0, 1.0, 1.0, 1.0};
#if defined(__CUDA_ARCH__)
    const double4 av = *reinterpret_cast<const double4 *>(q.a[k] + p);
    a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
    if (q.b[k]) { const double4 bv = *reinterpret_cast<const double4 *>(q.b[k] + p); b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w; }
#else
    for (int j = 0; j < 4; ++j) { a[j] = q.a[k][p + j]; if (q.b[k]) b[j] = q.b[k][p + j]; }
#endif
    const double sk = q.s[k];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fma(sk * a[j], b[j], v[j]);
  }
}

struct DevFFT {
  int nside, npair, nring;
  int Wn;                   // twiddle table size (power of two >= every M)
  const cplx *W;            // W[k] = e^{-2 pi i k / Wn}
  const cplx *V;            // Bluestein kernel spectra, DIF order, concatenated
  const long long *voff;    // [npair] offset into V (or -1)
  const int *M;             // [npair] FFT size: q (direct), nextpow2(2q-1) (Bluestein), 0 (tiny)
  const int *nphi;          // [npair]
  const int *shifted;       // [npair] 1: phi0 = pi/nphi
  const long long *start_n, *start_s;   // [npair] first pixel of the north / south ring (-1: none)
  const int *order;         // [npair] block -> ring pair, most expensive first
  const int *mtop;          // [npair] highest m the Legendre stage touches on this pair (or null = mmax)
  // m-partitioned analysis: column m of the phase array is written into the array of the rank that owns m
  int dist_n, dist_mblk;    // dist_n <= 1: single GPU
  cplx *dist_x[kMaxRanks];
  int nb4_maxm, nb1_minm;   // DFTs batched per pass: 4 (M <= nb4_maxm), 1 (M >= nb1_minm), else 2
  PixProg pix;              // analysis only: how a pixel value is obtained (pix.n == 0: read from the map argument)
};
PLK_HD int auto_nbatch(const DevFFT &f, int M) { return M == 0 ? 0 : (M >= f.nb1_minm ? 1 : (M <= f.nb4_maxm ? 4 : 2)); }
PLK_HD cplx *phase_out(const DevFFT &f, cplx *X, int m) {
  return f.dist_n > 1 ? f.dist_x[dist_owner_of_m(m, f.dist_mblk, f.dist_n)] : X;
}

PLK_HD int ilog2(int v) { int r = 0; while ((1 << r) < v) ++r; return r; }
PLK_HD int lg2(int v) {   // exact log2 of a power of two
#if defined(__CUDA_ARCH__)
  return 31 - __clz(v);
#else
  return 31 - __builtin_clz((unsigned)v);
#endif
}
PLK_HD int bitrev(int v, int bits) {
#if defined(__CUDA_ARCH__)
  return (int)(__brev((unsigned)v) >> (32 - bits));
#else
  unsigned r = 0;
  for (int i = 0; i < bits; ++i) { r = (r << 1) | ((unsigned)v & 1u); v >>= 1; }
  return (int)r;
#endif
}
// shared-memory element swizzle (bijective on every aligned block of 16 elements)
PLK_HD int SW(int i) { return i ^ ((i >> 4) & 15); }

// twiddle e^{sign * 2 pi i * num / L}, 0 <= num < L, from a QUARTER-wave table W[k] = e^{-2 pi i k / Wn}, k < Wn/4
// (kept in shared memory by the kernels): the other quadrants follow from W[k + Wn/4] = -i W[k].
// All sizes are powers of two and enter as log2 so that no integer division is generated.
template <int SIGN>
PLK_HD cplx tw(const cplx *W, int lgWn, int num, int lgL) {
  const int idx = num << (lgWn - lgL);
  const int quad = idx >> (lgWn - 2);
  cplx w = W[idx & ((1 << (lgWn - 2)) - 1)];
  if (quad & 1) w = mul_mi(w);
  if (quad & 2) w = mk(-w.x, -w.y);
  return SIGN < 0 ? w : conj(w);
}

// z * e^{SIGN 2 pi i k / 16}, k = 0..7 known at compile time after unrolling
template <int SIGN>
PLK_HD cplx mul_root16(cplx z, int k) {
  const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, r2 = 0.70710678118654752440;
  const double sg = SIGN < 0 ? -1.0 : 1.0;   // e^{SIGN i phi} = cos phi + i sg sin phi
  switch (k) {
    case 0: return z;
    case 4: return SIGN < 0 ? mul_mi(z) : mul_i(z);
    case 2: return mk(r2 * (z.x - sg * z.y), r2 * (z.y + sg * z.x));
    case 6: return mk(-r2 * (z.x + sg * z.y), r2 * (sg * z.x - z.y));
    case 1: return mk(c1 * z.x - sg * s1 * z.y, c1 * z.y + sg * s1 * z.x);
    case 3: return mk(s1 * z.x - sg * c1 * z.y, s1 * z.y + sg * c1 * z.x);
    case 5: return mk(-s1 * z.x - sg * c1 * z.y, -s1 * z.y + sg * c1 * z.x);
    default: return mk(-c1 * z.x - sg * s1 * z.y, -c1 * z.y + sg * s1 * z.x);   // 7
  }
}

// log2(R) fused radix-2 DIF levels on v[k] = x[base + k * L/R] (j = index inside the stride block): results go back
// to the same places and the overall transform stays in bit-reversed order whatever the radix schedule.
// Level t pairs (k, k+h), h = R >> (t+1), with twiddle W_L^{2^t j} * e^{SIGN 2 pi i kk / (2h)}.
template <int R, int SIGN>
PLK_HD void radix_dif_regs(cplx (&v)[R], const cplx *W, int lgWn, int j, int lgL) {
  constexpr int lgR = R == 16 ? 4 : (R == 8 ? 3 : 2);
#pragma unroll
  for (int t = 0; t < lgR; ++t) {
    const int h = R >> (t + 1);
    const cplx wt = tw<SIGN>(W, lgWn, j << t, lgL);
#pragma unroll
    for (int g = 0; g < R; g += 2 * h) {
#pragma unroll
      for (int kk = 0; kk < h; ++kk) {
        const cplx a = v[g + kk], b = v[g + kk + h];
        v[g + kk] = a + b;
        v[g + kk + h] = mul_root16<SIGN>(a - b, kk * (8 / h)) * wt;
      }
    }
  }
}
// the transpose: DIT levels in the opposite order
template <int R, int SIGN>
PLK_HD void radix_dit_regs(cplx (&v)[R], const cplx *W, int lgWn, int j, int lgL) {
  constexpr int lgR = R == 16 ? 4 : (R == 8 ? 3 : 2);
#pragma unroll
  for (int t = lgR - 1; t >= 0; --t) {
    const int h = R >> (t + 1);
    const cplx wt = tw<SIGN>(W, lgWn, j << t, lgL);
#pragma unroll
    for (int g = 0; g < R; g += 2 * h) {
#pragma unroll
      for (int kk = 0; kk < h; ++kk) {
        const cplx a = v[g + kk];
        const cplx b = mul_root16<SIGN>(v[g + kk + h] * wt, kk * (8 / h));
        v[g + kk] = a + b;
        v[g + kk + h] = a - b;
      }
    }
  }
}

// Execution context: on the device a thread block; on the host (tests/emul) a single "thread" that walks
// every strided loop serially -- all cross-thread traffic goes through shared buffers between sync() points, so
// the very same body code is valid for both.
struct BlockCtx {
  PLK_HD int tid() const {
#if defined(__CUDA_ARCH__)
    return threadIdx.x;
#else
    return 0;
#endif
  }
  PLK_HD int nthr() const {
#if defined(__CUDA_ARCH__)
    return blockDim.x;
#else
    return 1;
#endif
  }
  PLK_HD void sync() const {
#if defined(__CUDA_ARCH__)
    __syncthreads();
#endif
  }
};

// one pass over nb transforms of M points stored at buf + f * M (swizzled): sub-transform length L = 2^lgL
template <int R, int SIGN, bool DIT, class Ctx>
PLK_HD void fft_pass(Ctx ctx, cplx *buf, int nb, int M, int lgL, const cplx *W, int lgWn) {
  constexpr int lgR = R == 16 ? 4 : (R == 8 ? 3 : 2);
  const int lgM = lg2(M), lgS = lgL - lgR, lgB = lgM - lgR;
  for (int w = ctx.tid(); w < (nb << lgB); w += ctx.nthr()) {
    const int f = w >> lgB, b = w & ((1 << lgB) - 1);
    const int grp = b >> lgS, j = b & ((1 << lgS) - 1);
    cplx *p = buf + ((size_t)f << lgM);
    const int base = (grp << lgL) + j;
    cplx v[R];
#pragma unroll
    for (int k = 0; k < R; ++k) v[k] = p[SW(base + (k << lgS))];
    if (DIT) radix_dit_regs<R, SIGN>(v, W, lgWn, j, lgL);
    else radix_dif_regs<R, SIGN>(v, W, lgWn, j, lgL);
#pragma unroll
    for (int k = 0; k < R; ++k) p[SW(base + (k << lgS))] = v[k];
  }
}
template <int SIGN, bool DIT, class Ctx>
PLK_HD void fft_pass_r(Ctx ctx, int r, cplx *buf, int nb, int M, int lgL, const cplx *W, int lgWn) {
  if (r == 4) fft_pass<16, SIGN, DIT>(ctx, buf, nb, M, lgL, W, lgWn);
  else if (r == 3) fft_pass<8, SIGN, DIT>(ctx, buf, nb, M, lgL, W, lgWn);
  else fft_pass<4, SIGN, DIT>(ctx, buf, nb, M, lgL, W, lgWn);
  ctx.sync();
}
// radix schedule: ceil(bits/4) passes of 2..4 bits each, larger ones first (bits >= 4)
PLK_HD int pass_bits(int bits, int ipass) {
  const int np = (bits + 3) >> 2;
  const int base = bits / np, extra = bits - base * np;
  return base + (ipass < extra ? 1 : 0);
}
// natural order in -> bit-reversed order out (batched, in place)
template <int SIGN, class Ctx>
PLK_HD void fft_dif(Ctx ctx, cplx *buf, int nb, int M, const cplx *W) {
  const int bits = lg2(M), np = (bits + 3) >> 2;
  int lgL = bits;
  for (int ip = 0; ip < np; ++ip) {
    const int r = pass_bits(bits, ip);
    fft_pass_r<SIGN, false>(ctx, r, buf, nb, M, lgL, W, bits);
    lgL -= r;
  }
}
// bit-reversed order in -> natural order out: the same passes backwards
template <int SIGN, class Ctx>
PLK_HD void fft_dit(Ctx ctx, cplx *buf, int nb, int M, const cplx *W) {
  const int bits = lg2(M), np = (bits + 3) >> 2;
  int lgL = 0;
  for (int ip = np - 1; ip >= 0; --ip) {
    const int r = pass_bits(bits, ip);
    lgL += r;
    fft_pass_r<SIGN, true>(ctx, r, buf, nb, M, lgL, W, bits);
  }
}

// ------------------------------------------------------------------ fold / unfold helpers
// Folded Hermitian spectrum of one ring, WITHOUT the phi0 factor e^{i pi k/n}:
//   D(k) = sum_{j>=0} sg^j X[k + j n] + sum_{j>=1} sg^j conj(X[j n - k]),  sg = -1 if shifted else +1
PLK_HD cplx fold_bin(const cplx *X, int mmax, int n, int k, int shifted) {
  cplx acc = mk(0.0, 0.0);
  double sg = 1.0;
  const double flip = shifted ? -1.0 : 1.0;
  // m = 0 enters with its real part only (a_l0 of a real field; healpy ignores Im a_l0 as well)
  for (int m = k; m <= mmax; m += n) { acc = acc + sg * (m == 0 ? mk(X[0].x, 0.0) : X[m]); sg *= flip; }
  sg = flip;
  for (int m = n - k; m <= mmax; m += n) { acc = acc + sg * conj(X[m]); sg *= flip; }
  return acc;
}

// E^{(a)}_{k'} = e^{(2a)}_{k'} + i e^{(2a+1)}_{k'},  e^{(r)} = w_n^{r k'} sum_c i^{rc} d_{k'+qc},
// d_k = e^{i pi k/n * shifted} D(k);  both a = 0 and a = 1 from one gather
PLK_HD void synth_input2(const cplx *X, int mmax, int n, int q, int kp, int shifted, cplx &E0, cplx &E1) {
  const cplx g = expipi32(kp, n);             // e^{i pi k'/n}
  const cplx g2 = g * g;                      // w_n^{k'}
  cplx d[4];
  // e^{i pi (k'+qc)/n} = g * e^{i pi c/4}
  const double r2 = 0.70710678118654752440;
  const cplx c8[4] = {mk(1.0, 0.0), mk(r2, r2), mk(0.0, 1.0), mk(-r2, r2)};
  if (mmax <= q) {
    // No aliasing beyond the first mirror image (every ring with n >= 4 mmax: the equatorial belt and most of the
    // caps): the four folded bins reduce to X[k'], X[q] (k' = 0 only) and conj X[q - k'] -- two independent loads
    // issued back to back instead of eight data-dependent loops.
    const cplx zero = mk(0.0, 0.0);
    const int mm = q - kp;                                     // 1 .. q
    const cplx x0 = kp <= mmax ? X[kp] : zero;
    const cplx xm = mm <= mmax ? X[mm] : zero;
    const cplx D0 = kp == 0 ? mk(x0.x, 0.0) : x0;
    const cplx D1 = (kp == 0 && q <= mmax) ? xm : zero;        // k = q exists only for k' = 0 (then mm = q)
    const cplx D3 = shifted ? mk(-xm.x, xm.y) : conj(xm);      // sg * conj(X[q - k'])
    d[0] = shifted ? (g * c8[0]) * D0 : D0;
    d[1] = shifted ? (g * c8[1]) * D1 : D1;
    d[2] = zero;
    d[3] = shifted ? (g * c8[3]) * D3 : D3;
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      cplx D = fold_bin(X, mmax, n, kp + q * c, shifted);
      d[c] = shifted ? (g * c8[c]) * D : D;
    }
  }
  const cplx s02 = d[0] + d[2], s13 = d[1] + d[3], m02 = d[0] - d[2], m13 = mul_i(d[1] - d[3]);
  const cplx g4 = g2 * g2, g6 = g4 * g2;
  // r = 0: sum_c d_c ; r = 1: i^c ; r = 2: (-1)^c ; r = 3: (-i)^c, times w_n^{r k'}
  const cplx e0 = s02 + s13, e1 = g2 * (m02 + m13), e2 = g4 * (s02 - s13), e3 = g6 * (m02 - m13);
  E0 = e0 + mul_i(e1);
  E1 = e2 + mul_i(e3);
}

PLK_HD cplx chirp(int t, int q) { return expipi32(t * t, q); }   // b_t = e^{i pi t^2/q}, t < 2^15

// In-place inverse-sign length-q DFTs of nb buffers: Z_k = sum_t z_t e^{+2 pi i t k/q}; Z_k = fetchZ(buf, k, ...)
template <class Ctx>
PLK_HD void idft_q(Ctx ctx, cplx *buf, int nb, int q, int M, const cplx *W, const cplx *Vq) {
  if (M == q) {
    fft_dif<+1>(ctx, buf, nb, M, W);
  } else {
    fft_dif<-1>(ctx, buf, nb, M, W);
    const int lgM = lg2(M);
    for (int w = ctx.tid(); w < (nb << lgM); w += ctx.nthr()) {
      const int f = w >> lgM, i = w & (M - 1);
      cplx *p = buf + ((size_t)f << lgM);
      p[SW(i)] = p[SW(i)] * Vq[i];
    }
    ctx.sync();
    fft_dit<+1>(ctx, buf, nb, M, W);
  }
}
PLK_HD cplx fetchZ(const cplx *buf, int k, int q, int M, int bits) {
  if (M == q) return buf[SW(bitrev(k, bits))];
  return (1.0 / (double)M) * (chirp(k, q) * buf[SW(k)]);
}
// copies the quarter-wave twiddles of an M-point transform into shared memory: tws[k] = e^{-2 pi i k/M}, k < M/4
template <class Ctx>
PLK_HD void load_twiddles(Ctx ctx, cplx *tws, int M, const cplx *Wg, int Wng) {
  const int st = Wng / M;
  for (int k = ctx.tid(); k < (M >> 2); k += ctx.nthr()) tws[k] = Wg[(size_t)k * st];
  ctx.sync();
}

// ------------------------------------------------------------------ synthesis: X[ring][m] -> pixels
// smem: [M/4 twiddles][nbatch buffers of M]; nbatch = 2 (one ring: a = 0, 1) or 4 (both rings of the pair)
template <class Ctx>
PLK_HD void ring_synth_body(Ctx ctx, const DevFFT &f, int ip, const cplx *X, int pitch, int mmax_in, double *map,
                            cplx *smem, int nbatch) {
  const int n = f.nphi[ip], q = n >> 2, M = f.M[ip], shifted = f.shifted[ip];
  // rows of X are exactly zero above mtop (pairs the Legendre stage skips): do not fold them
  const int mmax = f.mtop ? (f.mtop[ip] < mmax_in ? f.mtop[ip] : mmax_in) : mmax_in;
  const int nhalf = f.start_s[ip] >= 0 ? 2 : 1;
  if (M == 0) {
    // tiny ring: x_j = X_0 + 2 Re sum_{m>0} X_m e^{i m phi_j}, phi_j = pi (shifted + 2j)/n
    for (int w = ctx.tid(); w < nhalf * n; w += ctx.nthr()) {
      const int half = w / n, j = w - half * n;
      const cplx *Xr = X + (size_t)(half == 0 ? ip : f.nring - 1 - ip) * pitch;
      double acc = Xr[0].x;
      for (int m = 1; m <= mmax; ++m) {
        const cplx e = expipi32(m * (shifted + 2 * j), n);
        acc += 2.0 * (Xr[m].x * e.x - Xr[m].y * e.y);
      }
      map[(half == 0 ? f.start_n[ip] : f.start_s[ip]) + j] = acc;
    }
    return;
  }
  const int bits = lg2(M);
  const cplx *Vq = (f.voff[ip] >= 0) ? f.V + f.voff[ip] : nullptr;
  cplx *tws = smem, *buf = smem + (M >> 2);
  load_twiddles(ctx, tws, M, f.W, f.Wn);
  if (nbatch == 1) {
    // one DFT at a time (M = 8192 at nside 4096: two buffers do not fit in shared memory); the fold is redone for
    // a = 1 -- cheaper than a global-memory stash, and these are the long rings where little aliasing happens
    for (int pass = 0; pass < 2 * nhalf; ++pass) {
      const int half = pass >> 1, a = pass & 1;
      const cplx *Xr = X + (size_t)(half == 0 ? ip : f.nring - 1 - ip) * pitch;
      for (int i = ctx.tid(); i < M; i += ctx.nthr()) {
        cplx e = mk(0.0, 0.0);
        if (i < q) {
          cplx e0, e1;
          synth_input2(Xr, mmax, n, q, i, shifted, e0, e1);
          e = a ? e1 : e0;
          if (M != q) e = e * chirp(i, q);
        }
        buf[SW(i)] = e;
      }
      ctx.sync();
      idft_q(ctx, buf, 1, q, M, tws, Vq);
      double *out0 = map + (half == 0 ? f.start_n[ip] : f.start_s[ip]) + 2 * a;
      for (int t = ctx.tid(); t < q; t += ctx.nthr()) {
        const cplx y = fetchZ(buf, t, q, M, bits);
#if defined(__CUDA_ARCH__)
        *reinterpret_cast<double2 *>(out0 + 4 * t) = make_double2(y.x, y.y);
#else
        out0[4 * t] = y.x; out0[4 * t + 1] = y.y;
#endif
      }
      ctx.sync();
    }
    return;
  }
  const int hstep = nbatch >= 4 ? 2 : 1;
  for (int h0 = 0; h0 < nhalf; h0 += hstep) {
    const int nh = (nhalf - h0) < hstep ? (nhalf - h0) : hstep;
    for (int w = ctx.tid(); w < (nh << bits); w += ctx.nthr()) {
      const int hh = w >> bits, i = w & (M - 1), half = h0 + hh;
      cplx e0 = mk(0.0, 0.0), e1 = mk(0.0, 0.0);
      if (i < q) {
        const cplx *Xr = X + (size_t)(half == 0 ? ip : f.nring - 1 - ip) * pitch;
        synth_input2(Xr, mmax, n, q, i, shifted, e0, e1);
        if (M != q) { const cplx c = chirp(i, q); e0 = e0 * c; e1 = e1 * c; }
      }
      buf[((size_t)(2 * hh) << bits) + SW(i)] = e0;
      buf[((size_t)(2 * hh + 1) << bits) + SW(i)] = e1;
    }
    ctx.sync();
    idft_q(ctx, buf, 2 * nh, q, M, tws, Vq);
    for (int w = ctx.tid(); w < nh * q; w += ctx.nthr()) {
      const int hh = w / q, t = w - hh * q, half = h0 + hh;
      const cplx y0 = fetchZ(buf + ((size_t)(2 * hh) << bits), t, q, M, bits);
      const cplx y1 = fetchZ(buf + ((size_t)(2 * hh + 1) << bits), t, q, M, bits);
      double *out = map + (half == 0 ? f.start_n[ip] : f.start_s[ip]) + 4 * t;   // 32-byte aligned
#if defined(__CUDA_ARCH__)
      *reinterpret_cast<double4 *>(out) = make_double4(y0.x, y0.y, y1.x, y1.y);
#else
      out[0] = y0.x; out[1] = y0.y; out[2] = y1.x; out[3] = y1.y;
#endif
    }
    ctx.sync();
  }
}

// ------------------------------------------------------------------ analysis: pixels -> X[ring][m]
// X_m = wgt * sum_j map_j e^{-i m phi_j}
template <class Ctx>
PLK_HD void ring_anal_body(Ctx ctx, const DevFFT &f, int ip, const double *map, cplx *X, int pitch, int mmax_in,
                           double wgt, cplx *smem, int nbatch) {
  const int n = f.nphi[ip], q = n >> 2, M = f.M[ip], shifted = f.shifted[ip];
  // the Legendre stage never reads X above mtop on this pair
  const int mmax = f.mtop ? (f.mtop[ip] < mmax_in ? f.mtop[ip] : mmax_in) : mmax_in;
  const int nhalf = f.start_s[ip] >= 0 ? 2 : 1;
  if (M == 0) {
    for (int w = ctx.tid(); w < nhalf * (mmax + 1); w += ctx.nthr()) {
      const int half = w / (mmax + 1), m = w - half * (mmax + 1);
      const long long p0 = half == 0 ? f.start_n[ip] : f.start_s[ip];
      cplx acc = mk(0.0, 0.0);
      for (int j = 0; j < n; ++j) acc = acc + pix_eval(f.pix, map, p0 + j) * expipi32(-m * (shifted + 2 * j), n);
      phase_out(f, X, m)[(size_t)(half == 0 ? ip : f.nring - 1 - ip) * pitch + m] = wgt * acc;
    }
    return;
  }
  const int bits = lg2(M);
  const cplx *Vq = (f.voff[ip] >= 0) ? f.V + f.voff[ip] : nullptr;
  cplx *tws = smem, *buf = smem + (M >> 2);
  load_twiddles(ctx, tws, M, f.W, f.Wn);
  if (nbatch == 1) {
    // one DFT at a time (see ring_synth_body): pass a = 0 writes S^(0) + w S^(1), pass a = 1 adds w^2 S^(2) + w^3 S^(3);
    // the same thread owns the same X element in both passes
    for (int pass = 0; pass < 2 * nhalf; ++pass) {
      const int half = pass >> 1, a = pass & 1;
      const long long p0 = (half == 0 ? f.start_n[ip] : f.start_s[ip]) + 2 * a;
      for (int i = ctx.tid(); i < M; i += ctx.nthr()) {
        cplx v = mk(0.0, 0.0);
        if (i < q) {
          v = mk(pix_eval(f.pix, map, p0 + 4 * i), -pix_eval(f.pix, map, p0 + 4 * i + 1));
          if (M != q) v = v * chirp(i, q);
        }
        buf[SW(i)] = v;
      }
      ctx.sync();
      idft_q(ctx, buf, 1, q, M, tws, Vq);
      const size_t xrow = (size_t)(half == 0 ? ip : f.nring - 1 - ip) * pitch;
      for (int m = ctx.tid(); m <= mmax; m += ctx.nthr()) {
        cplx *Xr = phase_out(f, X, m) + xrow;
        const int jn = m / n, k = m - jn * n;
        const int kp = k % q, kq = (q - kp) % q;
        const cplx Y = conj(fetchZ(buf, kp, q, M, bits)), Yc = fetchZ(buf, kq, q, M, bits);
        const cplx Se = 0.5 * (Y + Yc), So = mul_mi(0.5 * (Y - Yc));
        const cplx g = expipi32(-k, n);
        const cplx g2 = g * g;
        cplx D = Se + g2 * So;
        if (a) D = (g2 * g2) * D;
        cplx ph = mk(wgt, 0.0);
        if (shifted) ph = ((jn & 1) ? -wgt : wgt) * g;
        const cplx r = ph * D;
        Xr[m] = a ? Xr[m] + r : r;
      }
      ctx.sync();
    }
    return;
  }
  const int hstep = nbatch >= 4 ? 2 : 1;
  for (int h0 = 0; h0 < nhalf; h0 += hstep) {
    const int nh = (nhalf - h0) < hstep ? (nhalf - h0) : hstep;
    for (int w = ctx.tid(); w < (nh << bits); w += ctx.nthr()) {
      const int hh = w >> bits, i = w & (M - 1), half = h0 + hh;
      cplx v0 = mk(0.0, 0.0), v1 = mk(0.0, 0.0);
      if (i < q) {
        const long long p0 = (half == 0 ? f.start_n[ip] : f.start_s[ip]) + 4 * i;
        if (f.pix.n == 0) {
          const double *in = map + p0;
          v0 = mk(in[0], -in[1]);      // conj(y^{(0)}_t), y^{(a)}_t = x_{4t+2a} + i x_{4t+2a+1}
          v1 = mk(in[2], -in[3]);
        } else {
          double pv[4];
          pix_eval4(f.pix, p0, pv);
          v0 = mk(pv[0], -pv[1]);
          v1 = mk(pv[2], -pv[3]);
        }
        if (M != q) { const cplx c = chirp(i, q); v0 = v0 * c; v1 = v1 * c; }
      }
      buf[((size_t)(2 * hh) << bits) + SW(i)] = v0;
      buf[((size_t)(2 * hh + 1) << bits) + SW(i)] = v1;
    }
    ctx.sync();
    idft_q(ctx, buf, 2 * nh, q, M, tws, Vq);
    // Y^{(a)}_k = conj(Z^{(a)}_k);  S^{(2a)}_k = (Y_k + conj Y_{q-k})/2,  S^{(2a+1)}_k = (Y_k - conj Y_{q-k})/(2i)
    // D_k = sum_r w_n^{-r k} S^{(r)}_{k mod q} ;  X_m = wgt e^{-i pi m/n * shifted} D_{m mod n}
    for (int w = ctx.tid(); w < nh * (mmax + 1); w += ctx.nthr()) {
      const int hh = w / (mmax + 1), m = w - hh * (mmax + 1), half = h0 + hh;
      const int jn = m / n, k = m - jn * n;
      const int kp = k % q, kq = (q - kp) % q;
      const cplx *b0 = buf + ((size_t)(2 * hh) << bits), *b1 = buf + ((size_t)(2 * hh + 1) << bits);
      const cplx Y0 = conj(fetchZ(b0, kp, q, M, bits)), Y0c = fetchZ(b0, kq, q, M, bits);
      const cplx Y1 = conj(fetchZ(b1, kp, q, M, bits)), Y1c = fetchZ(b1, kq, q, M, bits);
      const cplx S0 = 0.5 * (Y0 + Y0c), S1 = mul_mi(0.5 * (Y0 - Y0c));
      const cplx S2 = 0.5 * (Y1 + Y1c), S3 = mul_mi(0.5 * (Y1 - Y1c));
      const cplx g = expipi32(-k, n);                        // e^{-i pi k/n}
      const cplx g2 = g * g, g4 = g2 * g2, g6 = g4 * g2;     // w_n^{-k}, w_n^{-2k}, w_n^{-3k}
      const cplx D = (S0 + g2 * S1) + (g4 * S2 + g6 * S3);
      cplx ph = mk(wgt, 0.0);
      if (shifted) ph = ((jn & 1) ? -wgt : wgt) * g;
      phase_out(f, X, m)[(size_t)(half == 0 ? ip : f.nring - 1 - ip) * pitch + m] = ph * D;
    }
    ctx.sync();
  }
}

// Bluestein kernel spectrum for ring pair ip: v[t mod M] = e^{-i pi t^2/q}, |t| < q, forward DIF (kept in DIF order)
template <class Ctx>
PLK_HD void bluestein_setup_body(Ctx ctx, const DevFFT &f, int ip, cplx *Vout, cplx *smem) {
  if (f.voff[ip] < 0) return;
  const int q = f.nphi[ip] >> 2, M = f.M[ip];
  cplx *tws = smem, *buf = smem + (M >> 2);
  load_twiddles(ctx, tws, M, f.W, f.Wn);
  for (int i = ctx.tid(); i < M; i += ctx.nthr()) {
    cplx v = mk(0.0, 0.0);
    if (i < q) v = conj(chirp(i, q));
    else if (i > M - q) v = conj(chirp(M - i, q));
    buf[SW(i)] = v;
  }
  ctx.sync();
  fft_dif<-1>(ctx, buf, 1, M, tws);
  for (int i = ctx.tid(); i < M; i += ctx.nthr()) Vout[f.voff[ip] + i] = buf[SW(i)];
}

#if defined(__CUDACC__)
// one launch per FFT size class: `list` holds the ring pairs of the class, dynamic smem = (nbatch + 1/4) M complex.
// <256, 3>: up to three resident blocks (<= 80 registers); <512, 1>: the largest transforms, one block per SM.
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
ring_synth_kernel(DevFFT f, const int *__restrict__ list, int nbatch, const cplx *__restrict__ X, int pitch, int mmax,
                  double *__restrict__ map) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ip = list[blockIdx.x];
  // nbatch == 0: merged launch over all size classes (small plans, where launch latency beats occupancy)
  ring_synth_body(BlockCtx(), f, ip, X, pitch, mmax, map, reinterpret_cast<cplx *>(smem_raw),
                  nbatch ? nbatch : auto_nbatch(f, f.M[ip]));
}
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
ring_anal_kernel(DevFFT f, const int *__restrict__ list, int nbatch, const double *__restrict__ map, cplx *__restrict__ X,
                 int pitch, int mmax, double wgt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ip = list[blockIdx.x];
  ring_anal_body(BlockCtx(), f, ip, map, X, pitch, mmax, wgt, reinterpret_cast<cplx *>(smem_raw),
                 nbatch ? nbatch : auto_nbatch(f, f.M[ip]));
}
__global__ void __launch_bounds__(kFftThreads) bluestein_setup_kernel(DevFFT f, cplx *Vout) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bluestein_setup_body(BlockCtx(), f, blockIdx.x, Vout, reinterpret_cast<cplx *>(smem_raw));
}
#endif  // __CUDACC__

}  // namespace plk
