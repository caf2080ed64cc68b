// Ring FFT stage of the HEALPix SHT for sm_100a: phase array X[ring][m]  <->  RING-ordered pixels.
//
// One thread block per iso-latitude ring PAIR (north ring + southern twin have the same length n = 4q and
// the same phi0).  A real ring of n = 4q samples is handled as two complex length-q DFTs:
//   x_{4t+r} = sum_{k'<q} e^{(r)}_{k'} w_q^{t k'},   e^{(r)}_{k'} = w_n^{r k'} sum_{c<4} i^{rc} d_{k'+qc},
//   packed   y^{(a)}_t = x_{4t+2a} + i x_{4t+2a+1},  a = 0, 1,
// where d_k is the length-n Hermitian spectrum obtained by folding (aliasing) all m <= mmax onto the ring's
// band, including the e^{i m phi0} shift.  The length-q DFT runs entirely in shared memory:
//   q = 2^j        : in-place radix-4 DIF, output read in bit-reversed order;
//   otherwise      : Bluestein chirp-z with M = nextpow2(2q-1): DIF forward, pointwise multiply by the
//                    precomputed kernel spectrum (stored in DIF output order), DIT inverse -- no bit reversal;
//   q <= kTinyQ    : direct evaluation of the defining sum.
// Analysis uses the same inverse-DFT machinery on conjugated data (DFT(y) = conj(IDFT(conj y))).
// No cuFFT: 2047 distinct cap-ring lengths per nside would mean thousands of plans and launches per transform.
#pragma once
#include <cuda_runtime.h>

#include "plk_common.h"

namespace plk {

constexpr int kTinyQ = 8;
constexpr int kFftThreads = 256;

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
};

PLK_HD int ilog2(int v) { int r = 0; while ((1 << r) < v) ++r; return r; }
PLK_HD int bitrev(int v, int bits) {
  unsigned r = 0;
  for (int i = 0; i < bits; ++i) { r = (r << 1) | ((unsigned)v & 1u); v >>= 1; }
  return (int)r;
}

PLK_HD int lg2(int v) {   // exact log2 of a power of two
#if defined(__CUDA_ARCH__)
  return 31 - __clz(v);
#else
  return 31 - __builtin_clz((unsigned)v);
#endif
}

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

// ------------------------------------------------------------------ in-place shared-memory FFT stages
// "tid/nthr" explicit so the very same code can be driven from a host loop in tests.
template <int SIGN>
PLK_HD void dif_r2_stage(cplx *u, int M, const cplx *W, int lgWn, int tid, int nthr) {
  const int H = M >> 1, lgM = lg2(M);
  for (int i = tid; i < H; i += nthr) {
    cplx a = u[i], b = u[i + H];
    u[i] = a + b;
    u[i + H] = (a - b) * tw<SIGN>(W, lgWn, i, lgM);
  }
}
// fused pair of radix-2 DIF stages on sub-transforms of length L (bit-reversal compatible ordering)
template <int SIGN>
PLK_HD void dif_r4_stage(cplx *u, int M, int L, const cplx *W, int lgWn, int tid, int nthr) {
  const int Q = L >> 2, lgL = lg2(L), lgQ = lgL - 2;
  for (int b = tid; b < (M >> 2); b += nthr) {
    const int grp = b >> lgQ, j = b & (Q - 1);
    cplx *p = u + ((size_t)grp << lgL) + j;
    cplx a0 = p[0], a1 = p[Q], a2 = p[2 * Q], a3 = p[3 * Q];
    cplx t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3, t3 = a1 - a3;
    t3 = SIGN < 0 ? mul_mi(t3) : mul_i(t3);
    p[0] = t0 + t2;
    p[Q] = (t0 - t2) * tw<SIGN>(W, lgWn, 2 * j, lgL);
    p[2 * Q] = (t1 + t3) * tw<SIGN>(W, lgWn, j, lgL);
    p[3 * Q] = (t1 - t3) * tw<SIGN>(W, lgWn, 3 * j, lgL);
  }
}
// fused pair of radix-2 DIT stages producing sub-transforms of length L (input: two levels of bit reversal below)
template <int SIGN>
PLK_HD void dit_r4_stage(cplx *u, int M, int L, const cplx *W, int lgWn, int tid, int nthr) {
  const int Q = L >> 2, lgL = lg2(L), lgQ = lgL - 2;
  for (int b = tid; b < (M >> 2); b += nthr) {
    const int grp = b >> lgQ, j = b & (Q - 1);
    cplx *p = u + ((size_t)grp << lgL) + j;
    cplx e0 = p[0], e1 = p[Q], e2 = p[2 * Q], e3 = p[3 * Q];
    const cplx w2 = tw<SIGN>(W, lgWn, 2 * j, lgL);
    cplx e1w = e1 * w2, e3w = e3 * w2;
    cplx f0 = e0 + e1w, f1 = e0 - e1w, f2 = e2 + e3w, f3 = e2 - e3w;
    const cplx w1 = tw<SIGN>(W, lgWn, j, lgL);
    cplx f2w = f2 * w1;
    cplx f3w = f3 * w1;
    f3w = SIGN < 0 ? mul_mi(f3w) : mul_i(f3w);   // w^{j + L/4}
    p[0] = f0 + f2w;
    p[2 * Q] = f0 - f2w;
    p[Q] = f1 + f3w;
    p[3 * Q] = f1 - f3w;
  }
}
template <int SIGN>
PLK_HD void dit_r2_stage(cplx *u, int M, const cplx *W, int lgWn, int tid, int nthr) {
  const int H = M >> 1, lgM = lg2(M);
  for (int i = tid; i < H; i += nthr) {
    cplx a = u[i], b = u[i + H] * tw<SIGN>(W, lgWn, i, lgM);
    u[i] = a + b;
    u[i + H] = a - b;
  }
}

// Execution context: on the device a thread block; on the host (tests/emul) a single "thread" that walks
// every strided loop serially -- all cross-thread traffic goes through `buf` between sync() points, so the
// very same body code is valid for both.
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

// natural order in -> bit-reversed order out
template <int SIGN, class Ctx>
PLK_HD void fft_dif(Ctx ctx, cplx *u, int M, const cplx *W, int Wn) {
  int L = M;
  const int lgWn = lg2(Wn);
  if (lg2(M) & 1) { dif_r2_stage<SIGN>(u, M, W, lgWn, ctx.tid(), ctx.nthr()); ctx.sync(); L >>= 1; }
  for (; L >= 4; L >>= 2) { dif_r4_stage<SIGN>(u, M, L, W, lgWn, ctx.tid(), ctx.nthr()); ctx.sync(); }
}
// bit-reversed order in -> natural order out
template <int SIGN, class Ctx>
PLK_HD void fft_dit(Ctx ctx, cplx *u, int M, const cplx *W, int Wn) {
  const bool odd = lg2(M) & 1;
  const int Ltop = odd ? (M >> 1) : M;
  const int lgWn = lg2(Wn);
  for (int L = 4; L <= Ltop; L <<= 2) { dit_r4_stage<SIGN>(u, M, L, W, lgWn, ctx.tid(), ctx.nthr()); ctx.sync(); }
  if (odd) { dit_r2_stage<SIGN>(u, M, W, lgWn, ctx.tid(), ctx.nthr()); ctx.sync(); }
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
// d_k = e^{i pi k/n * shifted} D(k)
PLK_HD cplx synth_input(const cplx *X, int mmax, int n, int q, int kp, int shifted, int a) {
  const cplx g = expipi32(kp, n);             // e^{i pi k'/n}
  const cplx g2 = g * g;                      // w_n^{k'}
  cplx d[4];
  // e^{i pi (k'+qc)/n} = g * e^{i pi c/4}
  const double r2 = 0.70710678118654752440;
  const cplx c8[4] = {mk(1.0, 0.0), mk(r2, r2), mk(0.0, 1.0), mk(-r2, r2)};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    cplx D = fold_bin(X, mmax, n, kp + q * c, shifted);
    d[c] = shifted ? (g * c8[c]) * D : D;
  }
  // r = 2a (even): i^{rc} = (-1)^{a c};  r = 2a+1: i^{rc} = i^{(2a+1)c}
  cplx wr = mk(1.0, 0.0);
  for (int i = 0; i < 2 * a; ++i) wr = wr * g2;      // w_n^{2a k'}
  cplx e0, e1;
  if (a == 0) {
    e0 = d[0] + d[1] + d[2] + d[3];
    e1 = (d[0] - d[2]) + mul_i(d[1] - d[3]);          // i^c
  } else {
    e0 = (d[0] + d[2]) - (d[1] + d[3]);               // (-1)^c
    e1 = (d[0] - d[2]) - mul_i(d[1] - d[3]);          // i^{3c} = (-i)^c
  }
  e0 = wr * e0;
  e1 = (wr * g2) * e1;
  return e0 + mul_i(e1);
}

PLK_HD cplx chirp(int t, int q) { return expipi32(t * t, q); }   // b_t = e^{i pi t^2/q}, t < 2^15

// In-place inverse-sign length-q DFT of buf[0..q): Z_k = sum_t z_t e^{+2 pi i t k/q}.
// On return Z_k = fetchZ(buf, k, ...).
template <class Ctx>
PLK_HD void idft_q(Ctx ctx, cplx *buf, int q, int M, const cplx *W, int Wn, const cplx *Vq) {
  if (M == q) {
    fft_dif<+1>(ctx, buf, M, W, Wn);
  } else {
    fft_dif<-1>(ctx, buf, M, W, Wn);
    for (int i = ctx.tid(); i < M; i += ctx.nthr()) buf[i] = buf[i] * Vq[i];
    ctx.sync();
    fft_dit<+1>(ctx, buf, M, W, Wn);
  }
}
// copies the quarter-wave twiddles of an M-point transform next to the work buffer: tws[k] = e^{-2 pi i k/M}, k < M/4
template <class Ctx>
PLK_HD void load_twiddles(Ctx ctx, cplx *tws, int M, const cplx *Wg, int Wng) {
  const int st = Wng / M;
  for (int k = ctx.tid(); k < (M >> 2); k += ctx.nthr()) tws[k] = Wg[(size_t)k * st];
  ctx.sync();
}
PLK_HD cplx fetchZ(const cplx *buf, int k, int q, int M, int bits) {
  if (M == q) return buf[bitrev(k, bits)];
  return (1.0 / (double)M) * (chirp(k, q) * buf[k]);
}

// ------------------------------------------------------------------ synthesis: X[ring][m] -> pixels
template <class Ctx>
PLK_HD void ring_synth_body(Ctx ctx, const DevFFT &f, int ip, const cplx *X, int pitch, int mmax_in, double *map,
                            cplx *buf) {
  const int n = f.nphi[ip], q = n >> 2, M = f.M[ip], shifted = f.shifted[ip];
  const int bits = ilog2(M > 0 ? M : 1);
  const cplx *Vq = (f.voff[ip] >= 0) ? f.V + f.voff[ip] : nullptr;
  // rows of X are exactly zero above mtop (pairs the Legendre stage skips): do not fold them
  const int mmax = f.mtop ? (f.mtop[ip] < mmax_in ? f.mtop[ip] : mmax_in) : mmax_in;
  cplx *tws = buf + M;
  if (M > 0) load_twiddles(ctx, tws, M, f.W, f.Wn);
  for (int half = 0; half < 2; ++half) {
    const long long start = half == 0 ? f.start_n[ip] : f.start_s[ip];
    if (start < 0) continue;
    const int ring = half == 0 ? ip : f.nring - 1 - ip;
    const cplx *Xr = X + (size_t)ring * pitch;
    double *out = map + start;
    if (M == 0) {
      // tiny ring: x_j = X_0 + 2 Re sum_{m>0} X_m e^{i m phi_j}, phi_j = pi (shifted + 2j)/n
      for (int j = ctx.tid(); j < n; j += ctx.nthr()) {
        double acc = Xr[0].x;
        for (int m = 1; m <= mmax; ++m) {
          cplx e = expipi32(m * (shifted + 2 * j), n);
          acc += 2.0 * (Xr[m].x * e.x - Xr[m].y * e.y);
        }
        out[j] = acc;
      }
      continue;
    }
    for (int a = 0; a < 2; ++a) {
      for (int i = ctx.tid(); i < M; i += ctx.nthr()) {
        cplx v = mk(0.0, 0.0);
        if (i < q) {
          v = synth_input(Xr, mmax, n, q, i, shifted, a);
          if (M != q) v = v * chirp(i, q);
        }
        buf[i] = v;
      }
      ctx.sync();
      idft_q(ctx, buf, q, M, tws, M, Vq);
      for (int t = ctx.tid(); t < q; t += ctx.nthr()) {
        cplx y = fetchZ(buf, t, q, M, bits);
        out[4 * t + 2 * a] = y.x;
        out[4 * t + 2 * a + 1] = y.y;
      }
      ctx.sync();
    }
  }
}

// ------------------------------------------------------------------ analysis: pixels -> X[ring][m]
// X_m = wgt * sum_j map_j e^{-i m phi_j}
template <class Ctx>
PLK_HD void ring_anal_body(Ctx ctx, const DevFFT &f, int ip, const double *map, cplx *X, int pitch, int mmax_in,
                           double wgt, cplx *buf) {
  const int n = f.nphi[ip], q = n >> 2, M = f.M[ip], shifted = f.shifted[ip];
  const int bits = ilog2(M > 0 ? M : 1);
  const cplx *Vq = (f.voff[ip] >= 0) ? f.V + f.voff[ip] : nullptr;
  // the Legendre stage never reads X above mtop on this pair
  const int mmax = f.mtop ? (f.mtop[ip] < mmax_in ? f.mtop[ip] : mmax_in) : mmax_in;
  cplx *tws = buf + M;
  if (M > 0) load_twiddles(ctx, tws, M, f.W, f.Wn);
  for (int half = 0; half < 2; ++half) {
    const long long start = half == 0 ? f.start_n[ip] : f.start_s[ip];
    if (start < 0) continue;
    const int ring = half == 0 ? ip : f.nring - 1 - ip;
    cplx *Xr = X + (size_t)ring * pitch;
    const double *in = map + start;
    if (M == 0) {
      for (int m = ctx.tid(); m <= mmax; m += ctx.nthr()) {
        cplx acc = mk(0.0, 0.0);
        for (int j = 0; j < n; ++j) {
          cplx e = expipi32(-m * (shifted + 2 * j), n);
          acc = acc + in[j] * e;
        }
        Xr[m] = wgt * acc;
      }
      continue;
    }
    for (int a = 0; a < 2; ++a) {
      for (int i = ctx.tid(); i < M; i += ctx.nthr()) {
        cplx v = mk(0.0, 0.0);
        if (i < q) {
          v = mk(in[4 * i + 2 * a], -in[4 * i + 2 * a + 1]);   // conj(y_t)
          if (M != q) v = v * chirp(i, q);
        }
        buf[i] = v;
      }
      ctx.sync();
      idft_q(ctx, buf, q, M, tws, M, Vq);
      // Y_k = conj(Z_k);  S^{(2a)}_k = (Y_k + conj Y_{q-k})/2,  S^{(2a+1)}_k = (Y_k - conj Y_{q-k})/(2i)
      // D_k = sum_r w_n^{-r k} S^{(r)}_{k mod q} ;  X_m = wgt e^{-i pi m/n * shifted} D_{m mod n}
      for (int m = ctx.tid(); m <= mmax; m += ctx.nthr()) {
        const int jn = m / n, k = m - jn * n;
        const int kp = k % q, kq = (q - kp) % q;
        const cplx Yk = conj(fetchZ(buf, kp, q, M, bits));
        const cplx Yc = fetchZ(buf, kq, q, M, bits);           // conj(Y_{q-k})
        const cplx S0 = 0.5 * (Yk + Yc);
        const cplx S1 = mul_mi(0.5 * (Yk - Yc));
        const cplx g = expipi32(-k, n);                        // e^{-i pi k/n}
        const cplx g2 = g * g;                                 // w_n^{-k}
        cplx wr = mk(1.0, 0.0);
        for (int i = 0; i < 2 * a; ++i) wr = wr * g2;          // w_n^{-2a k}
        cplx D = wr * S0 + (wr * g2) * S1;
        cplx ph = mk(wgt, 0.0);
        if (shifted) ph = ((jn & 1) ? -wgt : wgt) * g;
        cplx val = ph * D;
        if (a == 0) Xr[m] = val;
        else Xr[m] = Xr[m] + val;
      }
      ctx.sync();
    }
  }
}

// Bluestein kernel spectrum for ring pair ip: v[t mod M] = e^{-i pi t^2/q}, |t| < q, forward DIF (kept in DIF order)
template <class Ctx>
PLK_HD void bluestein_setup_body(Ctx ctx, const DevFFT &f, int ip, cplx *Vout, cplx *buf) {
  if (f.voff[ip] < 0) return;
  const int q = f.nphi[ip] >> 2, M = f.M[ip];
  for (int i = ctx.tid(); i < M; i += ctx.nthr()) {
    cplx v = mk(0.0, 0.0);
    if (i < q) v = conj(chirp(i, q));
    else if (i > M - q) v = conj(chirp(M - i, q));
    buf[i] = v;
  }
  ctx.sync();
  cplx *tws = buf + M;
  load_twiddles(ctx, tws, M, f.W, f.Wn);
  fft_dif<-1>(ctx, buf, M, tws, M);
  for (int i = ctx.tid(); i < M; i += ctx.nthr()) Vout[f.voff[ip] + i] = buf[i];
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(kFftThreads)
ring_synth_kernel(DevFFT f, const cplx *__restrict__ X, int pitch, int mmax, double *__restrict__ map) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ring_synth_body(BlockCtx(), f, f.order[blockIdx.x], X, pitch, mmax, map, reinterpret_cast<cplx *>(smem_raw));
}
__global__ void __launch_bounds__(kFftThreads)
ring_anal_kernel(DevFFT f, const double *__restrict__ map, cplx *__restrict__ X, int pitch, int mmax, double wgt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ring_anal_body(BlockCtx(), f, f.order[blockIdx.x], map, X, pitch, mmax, wgt, reinterpret_cast<cplx *>(smem_raw));
}
__global__ void __launch_bounds__(kFftThreads) bluestein_setup_kernel(DevFFT f, cplx *Vout) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bluestein_setup_body(BlockCtx(), f, blockIdx.x, Vout, reinterpret_cast<cplx *>(smem_raw));
}
#endif  // __CUDACC__

}  // namespace plk
