// Wigner small-d transforms on Gauss-Legendre nodes (reference: plancklens/wigners/wigners.f90:566-684, the Fortran
// behind utils_spin.wignerc -> qresp.get_response / nhl.get_nhl):
//   wignerpos  : xi(x_i)  = sum_l cl_l (2l+1)/(4 pi) d^l_{s1 s2}(x_i)
//   wignercoeff: cl_l     = 2 pi sum_i f_i d^l_{s1 s2}(x_i)
// d^l_{m m'} by the three-term recurrence in l (Varshalovich 4.8.2) started from the closed form at
// l0 = max(|m|, |m'|):  xi_{mm'} sqrt((a+b)! / (a! b!)) sin^a(theta/2) cos^b(theta/2),  a = |m - m'|, b = |m + m'|.
// One thread per node; spins are tiny (|s| <= 6), so plain doubles need no rescaling.  Not a hot path: the point is
// that normalisations and N0 biases no longer need the Fortran extension.
#pragma once
#include <cuda_runtime.h>

#include "plk_common.h"

namespace plk {

struct WigCoef {            // per l >= l0: d^{l+1} = (A_l x + B_l) d^l - C_l d^{l-1}
  const double *A, *B, *C;
  int l0, lmax;
  double seed;              // xi sqrt((a+b)!/(a! b!))
  int a, b;
};

PLK_D double wig_seed(const WigCoef &w, double x) {
  double r = w.seed;
  const double sh = sqrt(0.5 * (1.0 - x)), ch = sqrt(0.5 * (1.0 + x));
  for (int i = 0; i < w.a; ++i) r *= sh;
  for (int i = 0; i < w.b; ++i) r *= ch;
  return r;
}

__global__ void wignerpos_kernel(WigCoef w, const double *__restrict__ clw /* cl (2l+1)/4pi */, const double *__restrict__ x,
                                 int nx, double *__restrict__ xi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nx) return;
  const double xx = x[i];
  double pm = 0.0, p = wig_seed(w, xx), acc = 0.0;
  for (int l = w.l0; l <= w.lmax; ++l) {
    acc = fma(clw[l], p, acc);
    const double pn = fma(fma(w.A[l], xx, w.B[l]), p, -w.C[l] * pm);
    pm = p; p = pn;
  }
  xi[i] = acc;
}

__global__ void wig_scale_kernel(const double *__restrict__ cl, int lmax, double *__restrict__ clw) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l <= lmax) clw[l] = cl[l] * (2.0 * l + 1.0) * 0.07957747154594767;   // (2l+1) / 4 pi
}

constexpr int kWigChunk = 64;
// partial[block][l] = sum over the block's nodes of f_i d^l(x_i); summed in fixed order by wigner_finish_kernel
__global__ void __launch_bounds__(256) wignercoeff_kernel(WigCoef w, const double *__restrict__ f, const double *__restrict__ x,
                                                          int nx, double *__restrict__ partial, int pitch) {
  __shared__ double sh[8][kWigChunk];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double xx = i < nx ? x[i] : 0.0;
  const double fi = i < nx ? f[i] : 0.0;
  double pm = 0.0, p = i < nx ? wig_seed(w, xx) : 0.0;
  double *out = partial + (size_t)blockIdx.x * pitch;
  for (int l = threadIdx.x; l < w.l0 && l <= w.lmax; l += blockDim.x) out[l] = 0.0;
  for (int lb = w.l0; lb <= w.lmax; lb += kWigChunk) {
    const int n = min(kWigChunk, w.lmax - lb + 1);
    for (int k = 0; k < n; ++k) {
      const int l = lb + k;
      double v = fi * p;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) sh[warp][k] = v;
      const double pn = fma(fma(w.A[l], xx, w.B[l]), p, -w.C[l] * pm);
      pm = p; p = pn;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) s += sh[q][k];
      out[lb + k] = s;
    }
    __syncthreads();
  }
}
__global__ void wigner_finish_kernel(const double *__restrict__ partial, int nblk, int pitch, int lmax, double scale,
                                     double *__restrict__ cl) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l > lmax) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += partial[(size_t)b * pitch + l];
  cl[l] = scale * s;
}

}  // namespace plk
