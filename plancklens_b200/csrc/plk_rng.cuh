// Counter-based Gaussian random numbers on the device: Philox4x32-10 (Salmon et al. 2011, "Parallel random numbers: as
// easy as 1, 2, 3") + Box-Muller in FP64.  Replaces, for the synthetic skies of the throughput runs, the host-side numpy
// draws of the phase libraries (reference: plancklens/sims/phas.py:137-195 keeps numpy RNG states in sqlite so that a
// phase can be regenerated; a counter-based generator regenerates element i of stream s from (seed, s, i) alone).
//   counter = (i lo, i hi, stream lo, stream hi), key = (seed lo, seed hi); one call yields 4 x 32 bits = two 53-bit
//   uniforms u1, u2 in (0, 1) -> z0 = sqrt(-2 ln u1) cos(2 pi u2), z1 = sqrt(-2 ln u1) sin(2 pi u2) = out[2i], out[2i+1].
// HBM-bound: 8 bytes written per normal, no reads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "plk_common.h"

namespace plk {

struct Philox4 { uint32_t v[4]; };

PLK_HD Philox4 philox4x32_10(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}
// 53-bit uniform in (0, 1) from two 32-bit words
PLK_HD double u53(uint32_t hi, uint32_t lo) {
  const uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}

#if defined(__CUDACC__)
// out[0 .. n): unit normals; scale != 1 multiplies them; add != nullptr: out[i] = add[i] + scale z_i (noise on a map)
__global__ void randn_kernel(uint64_t seed, uint64_t stream, long long n, double scale, const double *__restrict__ add,
                             double *__restrict__ out) {
  const long long npair = (n + 1) >> 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npair; i += (long long)gridDim.x * blockDim.x) {
    const Philox4 p = philox4x32_10((uint64_t)i, stream, seed);
    const double u1 = u53(p.v[0], p.v[1]), u2 = u53(p.v[2], p.v[3]);
    const double r = scale * sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    const long long j = 2 * i;
    if (j + 1 < n) {
      double2 z = make_double2(r * c, r * s);
      if (add) { const double2 a = *reinterpret_cast<const double2 *>(add + j); z.x += a.x; z.y += a.y; }
      *reinterpret_cast<double2 *>(out + j) = z;
    } else {
      out[j] = (add ? add[j] : 0.0) + r * c;
    }
  }
}
// raw 32-bit words (tests: bit-exact comparison with the numpy restatement): out[4 i + k]
__global__ void philox_words_kernel(uint64_t seed, uint64_t stream, long long ncalls, uint32_t *__restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ncalls; i += (long long)gridDim.x * blockDim.x) {
    const Philox4 p = philox4x32_10((uint64_t)i, stream, seed);
    *reinterpret_cast<uint4 *>(out + 4 * i) = make_uint4(p.v[0], p.v[1], p.v[2], p.v[3]);
  }
}
// alm phases of a real field (reference recipe: sims/phas.py:162-168): a = (z0 + i z1) / sqrt 2, real N(0, 1) at m = 0.
// Element i of the alm array uses Philox call i: re = z0, im = z1.  The m = 0 block is the first lmax + 1 entries.
__global__ void randn_alm_kernel(uint64_t seed, uint64_t stream, long long nalm, int lmax, cplx *__restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nalm; i += (long long)gridDim.x * blockDim.x) {
    const Philox4 p = philox4x32_10((uint64_t)i, stream, seed);
    const double u1 = u53(p.v[0], p.v[1]), u2 = u53(p.v[2], p.v[3]);
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    const double is2 = 0.70710678118654752440;
    out[i] = (i <= lmax) ? mk(r * c, 0.0) : mk(is2 * r * c, is2 * r * s);
  }
}
#endif

}  // namespace plk
