// HBM-bound helper kernels: alm pre/post scaling around the Legendre stage, alm BLAS-1 and per-pixel passes.
#pragma once
#include <cuda_runtime.h>

#include "plk_common.h"
#include "plk_legendre.cuh"

namespace plk {

// ------------------------------------------------------------------ synthesis input records
// grid: (ceil((lmax+1)/256), mmax+1).  spin 0: rec = double2 alpha*fl1*a ; spin s: rec = double4 {H+, H-}
template <bool SPIN>
__global__ void prep_alm_kernel(DevSpin t, const cplx *__restrict__ a1, const cplx *__restrict__ a2,
                                const double *__restrict__ fl1, const double *__restrict__ fl2, void *__restrict__ rec,
                                const int *__restrict__ morder) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = morder[blockIdx.y];
  if (l > t.lmax || l < m) return;
  const int64_t i = alm_idx(t.lmax, l, m);
  const double al = t.alpha[i];
  const double s1 = al * (fl1 ? fl1[l] : 1.0);
  if (!SPIN) {
    const cplx a = a1[i];
    reinterpret_cast<double2 *>(rec)[i] = make_double2(s1 * a.x, s1 * a.y);
  } else {
    const double s2 = al * (fl2 ? fl2[l] : 1.0);
    const cplx g = a1[i];
    const cplx c = a2 ? a2[i] : mk(0.0, 0.0);
    const double gr = s1 * g.x, gi = s1 * g.y, cr = s2 * c.x, ci = s2 * c.y;
    const double sg = (t.spin & 1) ? -1.0 : 1.0;
    // H+ = -1/2 (G + iC) ; H- = -1/2 (-1)^s (G - iC)
    reinterpret_cast<double4 *>(rec)[i] =
        make_double4(-0.5 * (gr - ci), -0.5 * (gi + cr), -0.5 * sg * (gr + ci), -0.5 * sg * (gi - cr));
  }
}

// ------------------------------------------------------------------ analysis output
// sums the tile partials in fixed order and applies alpha, fl and the G/C recombination
template <bool SPIN>
__global__ void finish_alm_kernel(DevSpin t, const double *__restrict__ part, long long part_stride, int ntile,
                                  const double *__restrict__ fl1, const double *__restrict__ fl2,
                                  cplx *__restrict__ a1, cplx *__restrict__ a2, const int *__restrict__ morder,
                                  const cplx *__restrict__ add1 = nullptr, const double *__restrict__ afl1 = nullptr,
                                  const cplx *__restrict__ add2 = nullptr, const double *__restrict__ afl2 = nullptr) {
  // add1 / add2 (optional): out_c += afl_c[l] * add_c[l, m] -- the S^-1 x term of the CG forward operators
  // (opfilt_tt.py:67-73, opfilt_pp.py:51-55 with a diagonal S^-1) folded into the analysis output pass
  constexpr int NV = SPIN ? 4 : 2;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = morder[blockIdx.y];
  if (l > t.lmax || l < m) return;
  const int64_t i = alm_idx(t.lmax, l, m);
  const int l0 = m > t.spin ? m : t.spin;
  cplx e1 = mk(0.0, 0.0), e2 = mk(0.0, 0.0);
  if (add1) { const cplx x = add1[i]; const double f = afl1[l]; e1 = mk(f * x.x, f * x.y); }
  if (SPIN && add2) { const cplx x = add2[i]; const double f = afl2[l]; e2 = mk(f * x.x, f * x.y); }
  if (l < l0) {
    a1[i] = e1;
    if (SPIN) a2[i] = e2;
    return;
  }
  double v[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = 0.0;
  for (int tl = 0; tl < ntile; ++tl) {
    const double *p = part + (size_t)tl * part_stride + (size_t)i * NV;
    if (SPIN) {
      const double4 q = *reinterpret_cast<const double4 *>(p);
      v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w;
    } else {
      const double2 q = *reinterpret_cast<const double2 *>(p);
      v[0] += q.x; v[1] += q.y;
    }
  }
  const double al = t.alpha[i];
  const double s1 = al * (fl1 ? fl1[l] : 1.0);
  if (!SPIN) {
    a1[i] = mk(fma(s1, v[0], e1.x), fma(s1, v[1], e1.y));
  } else {
    const double s2 = al * (fl2 ? fl2[l] : 1.0);
    const double sg = (t.spin & 1) ? -1.0 : 1.0;
    // +a = S+, -a' = (-1)^s S-;  G = -1/2 (+a + -a') ; C = i/2 (+a - -a')
    const double pr = v[0], pi = v[1], mr = sg * v[2], mi = sg * v[3];
    a1[i] = mk(fma(-0.5 * s1, pr + mr, e1.x), fma(-0.5 * s1, pi + mi, e1.y));
    a2[i] = mk(fma(-0.5 * s2, pi - mi, e2.x), fma(0.5 * s2, pr - mr, e2.y));
  }
}

// ------------------------------------------------------------------ alm BLAS-1
__global__ void almxfl_kernel(int lmax, const cplx *__restrict__ in, const double *__restrict__ fl, int nfl,
                              cplx *__restrict__ out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (l > lmax || l < m) return;
  const int64_t i = alm_idx(lmax, l, m);
  const double f = l < nfl ? fl[l] : 0.0;
  const cplx a = in[i];
  out[i] = mk(f * a.x, f * a.y);
}

__global__ void axpy_kernel(long long n2, double a, const double *__restrict__ a_dev, const double *__restrict__ x,
                            double *__restrict__ y) {
  const double aa = a_dev ? *a_dev : a;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x)
    y[i] = fma(aa, x[i], y[i]);
}

PLK_D double block_sum(double v, double *sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = l < (blockDim.x >> 5) ? sh[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;   // valid in thread 0
}

// one block per m: partial[m] = w_m sum_{l>=max(m,lmin)} Re(a conj b)
__global__ void dot_partial_kernel(int lmax, int lmin, const cplx *__restrict__ a, const cplx *__restrict__ b,
                                   double *__restrict__ partial) {
  __shared__ double sh[32];
  const int m = blockIdx.x;
  double acc = 0.0;
  const int lo = m > lmin ? m : lmin;
  const int64_t base = alm_idx(lmax, 0, m);
  for (int l = lo + threadIdx.x; l <= lmax; l += blockDim.x) {
    const cplx x = a[base + l], y = b[base + l];
    acc = fma(x.x, y.x, fma(x.y, y.y, acc));
  }
  double r = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[m] = (m == 0 ? 1.0 : 2.0) * r;
}
// fixed-order final reduction, nout independent sums laid out as partial[j * n + i]
__global__ void final_sum_kernel(const double *__restrict__ partial, int n, int nout, double *__restrict__ out) {
  __shared__ double sh[32];
  for (int j = 0; j < nout; ++j) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[(size_t)j * n + i];
    double r = block_sum(acc, sh);
    if (threadIdx.x == 0) out[j] = r;
  }
}

// Last-block pattern: every block publishes its partial, takes a ticket; the block that draws the last ticket sums all
// partials in a FIXED order (thread-strided, then the block tree) -- the result does not depend on which block that is.
PLK_D bool last_block_ticket(unsigned int *counter, bool *s_last) {
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(counter, 1u);
    *s_last = (t == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (*s_last) __threadfence();
  return *s_last;
}
// One-kernel weighted dot product of up to four alm pairs (opfilt_tt.py:43-51, opfilt_pp.py:27-34, opfilt_tp.py:46-58)
// with the CG step-length arithmetic of cd_solve.py:69-71, :95-99 folded into the final sum:
//   out[0] = s = sum_j sum_{l >= lmin} w_m Re(a_j conj b_j)
//   num != null:  out[1] = scale * num[0] / s,  out[2] = -out[1]       (alpha = (d.r) / (d.Ad) and -alpha)
//   den != null:  out[1] = scale * s / den[0],  out[2] = -out[1]       (beta = -(d'.Ad) / (d.Ad))
// A zero divisor gives 0 (the reference's monitor stops a stage whose residual is exactly zero before dividing).
struct DotArgs { const cplx *a[4], *b[4]; int n; };
__global__ void dot_fused_kernel(DotArgs q, int lmax, int lmin, double *__restrict__ partial, unsigned int *counter,
                                 const double *__restrict__ num, const double *__restrict__ den, double scale,
                                 double *__restrict__ out) {
  __shared__ double sh[32];
  __shared__ bool s_last;
  const int j = blockIdx.x / (lmax + 1), m = blockIdx.x - j * (lmax + 1);
  const cplx *a = q.a[j], *b = q.b[j];
  double acc = 0.0;
  const int lo = m > lmin ? m : lmin;
  const int64_t base = alm_idx(lmax, 0, m);
  for (int l = lo + threadIdx.x; l <= lmax; l += blockDim.x) {
    const cplx x = a[base + l], y = b[base + l];
    acc = fma(x.x, y.x, fma(x.y, y.y, acc));
  }
  const double r = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = (m == 0 ? 1.0 : 2.0) * r;
  if (!last_block_ticket(counter, &s_last)) return;
  double tot = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) tot += __ldcg(partial + i);
  const double s = block_sum(tot, sh);
  if (threadIdx.x == 0) {
    out[0] = s;
    if (num) { const double v = s != 0.0 ? scale * (num[0] / s) : 0.0; out[1] = v; out[2] = -v; }
    else if (den) { const double d = den[0]; const double v = d != 0.0 ? scale * (s / d) : 0.0; out[1] = v; out[2] = -v; }
    *counter = 0u;
  }
}
// y1 += a x1 ; y2 -= a x2 (a from device memory): the solution and residual updates of one CG iteration
// (cd_solve.py:75, :82-84) in one pass
__global__ void axpy2_kernel(long long n2, const double *__restrict__ a_dev, const double *__restrict__ x1,
                             double *__restrict__ y1, const double *__restrict__ x2, double *__restrict__ y2) {
  const double aa = *a_dev;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    y1[i] = fma(aa, x1[i], y1[i]);
    y2[i] = fma(-aa, x2[i], y2[i]);
  }
}

// cl[l] = 1/(2l+1) sum_m w_m Re(a_lm conj b_lm), w_0 = 1, w_{m>0} = 2 (hp.alm2cl); one thread per l, coalesced over l
__global__ void alm2cl_kernel(int lmax, const cplx *__restrict__ a, const cplx *__restrict__ b, double *__restrict__ cl) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l > lmax) return;
  double acc = 0.0;
  for (int m = 0; m <= l; ++m) {
    const int64_t i = alm_idx(lmax, l, m);
    const cplx x = a[i], y = b[i];
    const double t = fma(x.x, y.x, x.y * y.y);
    acc += m == 0 ? t : 2.0 * t;
  }
  cl[l] = acc / (2.0 * l + 1.0);
}

// out[0] = scale * num[0] / den[0]: the CG step lengths (cd_solve.py:69-71, :95-99) without a host round trip
__global__ void scalar_ratio_kernel(const double *__restrict__ num, const double *__restrict__ den, double scale,
                                    double *__restrict__ out) {
  // a zero denominator gives 0: an exactly-zero residual (all-zero or fully masked input) must not poison the solve
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = den[0] != 0.0 ? scale * (num[0] / den[0]) : 0.0;
}

__global__ void alm_copy_kernel(int lmax_in, const cplx *__restrict__ in, int lmax_out, cplx *__restrict__ out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (l > lmax_out || l < m) return;
  out[alm_idx(lmax_out, l, m)] = (l <= lmax_in && m <= lmax_in) ? in[alm_idx(lmax_in, l, m)] : mk(0.0, 0.0);
}
__global__ void alm_splice_kernel(int lmax_lo, const cplx *__restrict__ lo, int lmax_hi, const cplx *__restrict__ hi,
                                  int lsplit, cplx *__restrict__ out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (l > lmax_hi || l < m) return;
  const int64_t i = alm_idx(lmax_hi, l, m);
  out[i] = (l <= lsplit) ? lo[alm_idx(lmax_lo, l, m)] : hi[i];
}

// out (lmax_hi) = lo for l <= lsplit, fl[l] * hi above: the diagonal high-l branch of multigrid.pre_op_split
// (multigrid.py:163-182 with `diag_cl`) and the splice in one pass -- same numbers as almxfl followed by alm_splice
__global__ void alm_splice_xfl_kernel(int lmax_lo, const cplx *__restrict__ lo, int lmax_hi, const cplx *__restrict__ hi,
                                      const double *__restrict__ fl, int nfl, int lsplit, cplx *__restrict__ out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (l > lmax_hi || l < m) return;
  const int64_t i = alm_idx(lmax_hi, l, m);
  if (l <= lsplit) { out[i] = lo[alm_idx(lmax_lo, l, m)]; return; }
  const double f = l < nfl ? fl[l] : 0.0;
  const cplx a = hi[i];
  out[i] = mk(f * a.x, f * a.y);
}

// ------------------------------------------------------------------ per-pixel passes
// partial[b] = sum over the block's grid-stride share of a_p b_p (fixed grid -> fixed summation order)
__global__ void map_dot_partial_kernel(long long n, const double *__restrict__ a, const double *__restrict__ b,
                                       double *__restrict__ partial) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc = fma(a[i], b[i], acc);
  const double r = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}
__global__ void map_mul_kernel(long long n, double *__restrict__ y, const double *__restrict__ a) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] *= a[i];
}
__global__ void map_mul2_kernel(long long n, double *__restrict__ g, double *__restrict__ c, const double *__restrict__ t) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double tt = t[i];
    g[i] *= tt; c[i] *= tt;
  }
}
__global__ void map_qe_pp_kernel(long long n, const double *__restrict__ q, const double *__restrict__ u,
                                 const double *__restrict__ g3, const double *__restrict__ c3,
                                 const double *__restrict__ g1, const double *__restrict__ c1, double *__restrict__ re,
                                 double *__restrict__ im) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double Q = q[i], U = u[i];
    // (Q - iU)(G3 + iC3) - (Q + iU)(G1 - iC1)
    const double r = (Q * g3[i] + U * c3[i]) - (Q * g1[i] + U * c1[i]);
    const double s = (Q * c3[i] - U * g3[i]) - (U * g1[i] - Q * c1[i]);
    re[i] = r; im[i] = s;
  }
}
__global__ void map_ninv3_kernel(long long n, double *__restrict__ q, double *__restrict__ u,
                                 const double *__restrict__ nqq, const double *__restrict__ nqu,
                                 const double *__restrict__ nuu) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double Q = q[i], U = u[i];
    q[i] = nqq[i] * Q + nqu[i] * U;
    u[i] = nuu[i] * U + nqu[i] * Q;
  }
}

// d += a * b for complex maps stored as (re, im) pairs of real maps; b_im may be null (real second factor)
__global__ void map_cmul_acc_kernel(long long n, const double *__restrict__ ar, const double *__restrict__ ai,
                                    const double *__restrict__ br, const double *__restrict__ bi,
                                    double *__restrict__ dr, double *__restrict__ di) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double a = ar[i], b = ai ? ai[i] : 0.0, c = br[i], d = bi ? bi[i] : 0.0;
    dr[i] += a * c - b * d;
    di[i] += a * d + b * c;
  }
}

// monopole / dipole templates: one block per ring; pixel direction from the ring geometry
struct DevRings {
  int nring;
  const long long *start;   // [nring]
  const int *nphi;          // [nring]
  const int *shifted;       // [nring]
  const double *z, *sth;    // [nring]
};
__global__ void modes_dot_kernel(DevRings r, double *__restrict__ m, const double *__restrict__ w,
                                 double *__restrict__ partial, unsigned int *counter, double *__restrict__ sums) {
  __shared__ double sh[32];
  __shared__ bool s_last;
  const int ir = blockIdx.x;
  const int n = r.nphi[ir];
  const long long st = r.start[ir];
  const double z = r.z[ir], s = r.sth[ir];
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    double v = m[st + j];
    if (w) { v *= w[st + j]; m[st + j] = v; }
    const cplx e = expipi_frac(r.shifted[ir] + 2 * j, n);
    a0 += v; a1 = fma(v, s * e.x, a1); a2 = fma(v, s * e.y, a2); a3 = fma(v, z, a3);
  }
  double r0 = block_sum(a0, sh), r1 = block_sum(a1, sh), r2 = block_sum(a2, sh), r3 = block_sum(a3, sh);
  if (threadIdx.x == 0) {
    partial[0 * r.nring + ir] = r0; partial[1 * r.nring + ir] = r1;
    partial[2 * r.nring + ir] = r2; partial[3 * r.nring + ir] = r3;
  }
  // the block with the last ticket sums the ring partials in ring order (fixed, deterministic)
  if (!last_block_ticket(counter, &s_last)) return;
  for (int j = 0; j < 4; ++j) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < r.nring; i += blockDim.x) acc += __ldcg(partial + (size_t)j * r.nring + i);
    const double t = block_sum(acc, sh);
    if (threadIdx.x == 0) sums[j] = t;
  }
  if (threadIdx.x == 0) *counter = 0u;
}
// m_p -= w_p * sum_a mode_a(p) coef_a,  coef = pinv(4x4, row-major) @ sums
__global__ void modes_sub_kernel(DevRings r, double *__restrict__ m, const double *__restrict__ w,
                                 const double *__restrict__ sums, const double *__restrict__ pinv) {
  const int ir = blockIdx.x;
  const int n = r.nphi[ir];
  const long long st = r.start[ir];
  const double z = r.z[ir], s = r.sth[ir];
  double c[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    c[a] = 0.0;
#pragma unroll
    for (int b = 0; b < 4; ++b) c[a] = fma(pinv[a * 4 + b], sums[b], c[a]);
  }
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const cplx e = expipi_frac(r.shifted[ir] + 2 * j, n);
    const double pm = c[0] + c[1] * (s * e.x) + c[2] * (s * e.y) + c[3] * z;
    const double ww = w ? w[st + j] : 1.0;
    m[st + j] -= ww * pm;
  }
}

// out = ca * a + cb * b on 2n doubles (b may be null: out = ca * a)
__global__ void lincomb_kernel(long long n2, double ca, const double *__restrict__ a, double cb,
                               const double *__restrict__ b, double *__restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x)
    out[i] = b ? fma(cb, b[i], ca * a[i]) : ca * a[i];
}

// out[l,m] = sum_j fl_j[l] * in_j[l,m], j < nterm <= 4 (Wiener-filter style combinations: C^EE E + C^TE T ...)
struct AlmTerms {
  const cplx *in[4];
  const double *fl[4];
  int nfl[4];
  int nterm;
};
__global__ void alm_combine_kernel(int lmax, AlmTerms t, cplx *__restrict__ out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (l > lmax || l < m) return;
  const int64_t i = alm_idx(lmax, l, m);
  double re = 0.0, im = 0.0;
  for (int j = 0; j < t.nterm; ++j) {
    const double f = l < t.nfl[j] ? t.fl[j][l] : 0.0;
    const cplx a = t.in[j][i];
    re = fma(f, a.x, re); im = fma(f, a.y, im);
  }
  out[i] = mk(re, im);
}

// X[r][m] = T[m][r] for rings r0 <= r < r1 and 0 <= m <= mmax (m-partitioned synthesis: exchange buffer -> phase array)
__global__ void phase_transpose_kernel(const cplx *__restrict__ T, int tpitch, cplx *__restrict__ X, int pitch, int mmax,
                                       int r0, int r1) {
  __shared__ cplx tile[32][33];
  const int rb = r0 + blockIdx.x * 32, mb = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int m = mb + i, r = rb + threadIdx.x;
    if (m <= mmax && r < r1) tile[i][threadIdx.x] = T[(size_t)m * tpitch + r];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = rb + i, m = mb + threadIdx.x;
    if (m <= mmax && r < r1) X[(size_t)r * pitch + m] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------ HEALPix degrade (sum of children), RING in / RING out
// hp.ud_grade(n_inv, nside_out, power=-2) of opfilt_tt.py:172-181 / opfilt_pp.py:244-251 without the RING <-> NEST
// permutations of the full-resolution map: one thread per OUTPUT pixel finds its (face, x, y) from the ring geometry,
// walks its (nside_in / nside_out)^2 children in a fixed order and looks each one up at its RING position.
// Index arithmetic: the published HEALPix pixelisation (Gorski et al. 2005), faces 0-3 north, 4-7 equatorial, 8-11 south.
struct HpxFace { int jrll[12], jpll[12]; };
PLK_HD HpxFace hpx_faces() {
  HpxFace f = {{2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4}, {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7}};
  return f;
}
PLK_HD long long hpx_isqrt(long long v) {
  long long r = (long long)sqrt((double)v + 0.5);
  while (r * r > v) --r;
  while ((r + 1) * (r + 1) <= v) ++r;
  return r;
}
PLK_HD void hpx_ring2xyf(long long nside, long long pix, int &ix, int &iy, int &face) {
  const HpxFace F = hpx_faces();
  const long long ncap = 2 * nside * (nside - 1), npix = 12 * nside * nside, nl2 = 2 * nside, nl4 = 4 * nside;
  long long iring, iphi, kshift, nr;
  if (pix < ncap) {
    iring = (1 + hpx_isqrt(1 + 2 * pix)) >> 1;
    iphi = pix + 1 - 2 * iring * (iring - 1);
    kshift = 0; nr = iring;
    face = (int)((iphi - 1) / nr);
  } else if (pix < npix - ncap) {
    const long long ip = pix - ncap, tmp = ip / nl4;
    iring = tmp + nside;
    iphi = ip - tmp * nl4 + 1;
    kshift = (iring + nside) & 1;
    nr = nside;
    const long long ire = tmp + 1, irm = nl2 + 1 - tmp;
    const long long ifm = (iphi - ire / 2 + nside - 1) / nside, ifp = (iphi - irm / 2 + nside - 1) / nside;
    face = (int)((ifp == ifm) ? (ifp | 4) : ((ifp < ifm) ? ifp : (ifm + 8)));
  } else {
    const long long ip = npix - pix;
    iring = (1 + hpx_isqrt(2 * ip - 1)) >> 1;
    iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
    kshift = 0; nr = iring;
    iring = 2 * nl2 - iring;
    face = 8 + (int)((iphi - 1) / nr);
  }
  const long long irt = iring - (long long)F.jrll[face] * nside + 1;
  long long ipt = 2 * iphi - (long long)F.jpll[face] * nr - kshift - 1;
  if (ipt >= nl2) ipt -= 8 * nside;
  ix = (int)((ipt - irt) >> 1);
  iy = (int)((-ipt - irt) >> 1);
}
PLK_HD long long hpx_xyf2ring(long long nside, int ix, int iy, int face) {
  const HpxFace F = hpx_faces();
  const long long nl4 = 4 * nside, ncap = 2 * nside * (nside - 1), npix = 12 * nside * nside;
  const long long jr = (long long)F.jrll[face] * nside - ix - iy - 1;
  long long nr, n_before, kshift;
  if (jr < nside) { nr = jr; n_before = 2 * nr * (nr - 1); kshift = 0; }
  else if (jr > 3 * nside) { nr = nl4 - jr; n_before = npix - 2 * (nr + 1) * nr; kshift = 0; }
  else { nr = nside; n_before = ncap + (jr - nside) * nl4; kshift = (jr - nside) & 1; }
  long long jp = ((long long)F.jpll[face] * nr + ix - iy + 1 + kshift) / 2;
  if (jp > nl4) jp -= nl4;
  else if (jp < 1) jp += nl4;
  return n_before + jp - 1;
}
// Sum of the f^2 children of output pixel (face, ix, iy) in the ORDER numpy uses for
// `nest_map.reshape(npix_out, f * f).sum(axis=1)` -- what healpy.ud_grade and the oracle do: children in NEST order
// (x in the even bits of the child index, y in the odd ones), fewer than 8 terms added one after the other, otherwise
// numpy's pairwise summation (eight interleaved partial sums per block of <= 128 terms, combined as a balanced tree;
// longer rows are halved recursively).  Matching the order keeps the degraded inverse-noise maps bit-identical to the
// reference's, which matters more than it should: the dense coarse preconditioner inverts "all but the ntmpl lowest
// eigenmodes" (dense.py:94-105), and when eigenvalues around that cut are nearly degenerate a last-bit change of the
// matrix swaps the modes and moves the whole CG trajectory by ~eps_min.
PLK_HD int hpx_compress_even_bits(unsigned int v) {
  v &= 0x55555555u;
  v = (v | (v >> 1)) & 0x33333333u;
  v = (v | (v >> 2)) & 0x0F0F0F0Fu;
  v = (v | (v >> 4)) & 0x00FF00FFu;
  v = (v | (v >> 8)) & 0x0000FFFFu;
  return (int)v;
}
PLK_HD double hpx_child(const double *in, int nside_in, int f, int ix, int iy, int face, int k) {
  const int dx = hpx_compress_even_bits((unsigned int)k), dy = hpx_compress_even_bits((unsigned int)k >> 1);
  return in[hpx_xyf2ring(nside_in, ix * f + dx, iy * f + dy, face)];
}
PLK_HD double hpx_block_sum(const double *in, int nside_in, int f, int ix, int iy, int face, int k0, int n) {   // 8 <= n <= 128
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = hpx_child(in, nside_in, f, ix, iy, face, k0 + j);
  for (int i = 8; i < n; i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] += hpx_child(in, nside_in, f, ix, iy, face, k0 + i + j);
  }
  return ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
}
PLK_HD double hpx_children_sum(const double *in, int nside_in, int nside_out, int ix, int iy, int face) {
  const int f = nside_in / nside_out, fac = f * f;
  if (fac < 8) {
    double res = 0.0;
    for (int k = 0; k < fac; ++k) res += hpx_child(in, nside_in, f, ix, iy, face, k);
    return res;
  }
  if (fac <= 128) return hpx_block_sum(in, nside_in, f, ix, iy, face, 0, fac);
  // fac = 4^j > 128: numpy halves the row until the pieces hold 128 terms; the halves are added as a balanced tree
  double blk[128];
  int nb = fac / 128;
  if (nb > 128) nb = 128;                      // f <= 128 (nside ratio); larger ratios fall back to longer leaves below
  const int leaf = fac / nb;
  for (int b = 0; b < nb; ++b) {
    if (leaf <= 128) blk[b] = hpx_block_sum(in, nside_in, f, ix, iy, face, b * leaf, leaf);
    else { double t = 0.0; for (int k = 0; k < leaf; ++k) t += hpx_child(in, nside_in, f, ix, iy, face, b * leaf + k); blk[b] = t; }
  }
  for (int n = nb; n > 1; n >>= 1)
    for (int b = 0; b < n / 2; ++b) blk[b] = blk[2 * b] + blk[2 * b + 1];
  return blk[0];
}
__global__ void udgrade_sum_kernel(int nside_in, const double *__restrict__ in, int nside_out, double *__restrict__ out) {
  const long long npo = 12LL * nside_out * nside_out;
  for (long long po = (long long)blockIdx.x * blockDim.x + threadIdx.x; po < npo; po += (long long)gridDim.x * blockDim.x) {
    int ix, iy, face;
    hpx_ring2xyf(nside_out, po, ix, iy, face);
    out[po] = hpx_children_sum(in, nside_in, nside_out, ix, iy, face);
  }
}

// real-harmonic packing used by the dense preconditioner (reference: qcinv/dense.py:16-53)
// lmax_src >= lmax: layout of the alm array read from (the low-l block of a longer vector is packed without a copy)
__global__ void alm2rlm_kernel(int lmax, const cplx *__restrict__ alm, double *__restrict__ rlm, int lmax_src) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (l > lmax || l < m) return;
  const cplx a = alm[alm_idx(lmax_src, l, m)];
  const double rt2 = 1.4142135623730951;
  if (m == 0) rlm[(size_t)l * l] = a.x;
  else { rlm[(size_t)l * l + 2 * m - 1] = a.x * rt2; rlm[(size_t)l * l + 2 * m] = a.y * rt2; }
}
__global__ void rlm2alm_kernel(int lmax, const double *__restrict__ rlm, cplx *__restrict__ alm) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (l > lmax || l < m) return;
  const double ir2 = 0.70710678118654757;
  alm[alm_idx(lmax, l, m)] = (m == 0) ? mk(rlm[(size_t)l * l], 0.0)
                                      : mk(rlm[(size_t)l * l + 2 * m - 1] * ir2, rlm[(size_t)l * l + 2 * m] * ir2);
}
// y = A x, A row-major n x n; one warp per row, fixed summation order
__global__ void matvec_kernel(int n, const double *__restrict__ A, const double *__restrict__ x, double *__restrict__ y) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const double *a = A + (size_t)row * n;
  double acc = 0.0;
  for (int j = lane; j < n; j += 32) acc = fma(a[j], x[j], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[row] = acc;
}

}  // namespace plk
