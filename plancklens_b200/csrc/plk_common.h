// Shared helpers for the B200 spin-weighted SHT library (host + device).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>

#if defined(__CUDACC__)
#define PLK_HD __host__ __device__ __forceinline__
#define PLK_D __device__ __forceinline__
#else
#define PLK_HD inline
#define PLK_D inline
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace plk {

struct cplx {
  double x, y;
};
PLK_HD cplx mk(double a, double b) { cplx r; r.x = a; r.y = b; return r; }
PLK_HD cplx operator+(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
PLK_HD cplx operator-(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
PLK_HD cplx operator*(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
PLK_HD cplx operator*(double s, cplx a) { return mk(s * a.x, s * a.y); }
PLK_HD cplx conj(cplx a) { return mk(a.x, -a.y); }
PLK_HD cplx mul_i(cplx a) { return mk(-a.y, a.x); }    // i*a
PLK_HD cplx mul_mi(cplx a) { return mk(a.y, -a.x); }   // -i*a

// ---- m-partitioned transforms over the GPUs of one NVSwitch box -------------------------
// The Legendre stage is split by m, the ring-FFT / pixel stage by ring pair.  m is dealt to ranks in blocks of
// `mblk` columns, boustrophedon (0 1 .. N-1 N-1 .. 1 0 ...) so that every rank gets the same mix of short (high m)
// and long (low m) recurrences; ring pairs are owned in contiguous blocks of (almost) equal pixel count.
constexpr int kMaxRanks = 8;
PLK_HD int dist_owner_of_m(int m, int mblk, int nranks) {
  const int c = (m / mblk) % (2 * nranks);
  return c < nranks ? c : 2 * nranks - 1 - c;
}
struct cplx;
struct DistX {                    // phase-array rows of ring pair ip live on rank q: pair_lo[q] <= ip < pair_lo[q+1]
  int nranks;                     // <= 1: single GPU (the plain X1 / X2 kernel arguments are used)
  int mblk;
  int tpitch;                     // row pitch of the TRANSPOSED exchange buffers T[m][ring] (synthesis)
  int pair_lo[kMaxRanks + 1];
  cplx *x1[kMaxRanks], *x2[kMaxRanks];   // every rank's exchange buffers (own + CUDA-IPC peer mappings over NVLink)
};

// healpy m-major triangular index (mmax == lmax): idx(l,m) = m(2 lmax + 1 - m)/2 + l
PLK_HD int64_t alm_idx(int lmax, int l, int m) { return (int64_t)m * (2 * lmax + 1 - m) / 2 + l; }
PLK_HD int64_t alm_size(int lmax, int mmax) { return (int64_t)mmax * (2 * lmax + 1 - mmax) / 2 + lmax + 1; }

// e^{i pi a / n} for integers a (any sign), n > 0, with exact argument reduction.
PLK_HD cplx expipi_frac(int64_t a, int64_t n) {
  int64_t two_n = 2 * n;
  int64_t r = a % two_n;
  if (r < 0) r += two_n;
  double s, c;
#if defined(__CUDA_ARCH__)
  sincospi((double)r / (double)n, &s, &c);
#else
  double t = (double)r / (double)n;
  // reduce to [-0.25, 0.25] turns of pi for accuracy on the host too
  int oct = (int)floor(t * 2.0 + 0.5);     // nearest multiple of 1/2
  double f = t - 0.5 * oct;                // |f| <= 0.25
  double sf = sin(M_PI * f), cf = cos(M_PI * f);
  switch (oct & 3) {
    case 0: s = sf; c = cf; break;
    case 1: s = cf; c = -sf; break;
    case 2: s = -sf; c = -cf; break;
    default: s = -cf; c = sf; break;
  }
#endif
  return mk(c, s);
}

// same with 32-bit integer arithmetic (|a| < 2^31, n < 2^30): the form the ring kernels use
PLK_HD cplx expipi32(int a, int n) {
  const int two_n = 2 * n;
  int r = a % two_n;
  if (r < 0) r += two_n;
#if defined(__CUDA_ARCH__)
  double s, c;
  sincospi((double)r / (double)n, &s, &c);
  return mk(c, s);
#else
  return expipi_frac(r, n);
#endif
}

// ---------------------------------------------------------------- extended-range double-double (setup only)
// value = (hi + lo) * 2^e, hi normalised to [0.5, 1)
struct xdd {
  double hi, lo;
  int e;
};
PLK_HD xdd xdd_norm(double hi, double lo, int e) {
  xdd r;
  if (hi == 0.0) { r.hi = 0; r.lo = 0; r.e = 0; return r; }
  int k;
  double h = frexp(hi, &k);
  r.hi = h; r.lo = ldexp(lo, -k); r.e = e + k;
  return r;
}
PLK_HD xdd xdd_mul(xdd a, xdd b) {
  double p = a.hi * b.hi;
  double err = fma(a.hi, b.hi, -p);
  double lo = err + (a.hi * b.lo + a.lo * b.hi);
  double s = p + lo;
  double l2 = lo - (s - p);
  return xdd_norm(s, l2, a.e + b.e);
}
PLK_HD xdd xdd_pow(xdd base, int n) {   // n >= 0
  xdd r; r.hi = 0.5; r.lo = 0; r.e = 1;  // 1.0
  while (n > 0) {
    if (n & 1) r = xdd_mul(r, base);
    n >>= 1;
    if (n) base = xdd_mul(base, base);
  }
  return r;
}

}  // namespace plk
