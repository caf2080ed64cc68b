// Host emulation of the CUDA kernels' arithmetic (TEST INFRASTRUCTURE ONLY, never loaded by the product).
// Drives the very same PLK_HD code (seed_one, fold/unfold, FFT stage functions, ring bodies) from plain host
// loops so that the algorithm can be checked against the oracle on a machine without a GPU.  The Legendre main
// loops are restated per (m, ring pair) with the kernel's exact operation order (tables, seeds, H+/H- records,
// even/odd sign handling); what is NOT covered here is the CUDA plumbing (TMA pipeline, warp butterflies).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../plancklens_b200/csrc/plk_common.h"
#include "../../plancklens_b200/csrc/plk_fft.cuh"
#include "../../plancklens_b200/csrc/plk_legendre.cuh"
#include "../../plancklens_b200/csrc/plk_tables.h"

using namespace plk;

namespace {
int nextpow2(int v) { int r = 1; while (r < v) r <<= 1; return r; }

struct HostFFT {
  DevFFT f;
  std::vector<int> M, nphi, shifted, order;
  std::vector<long long> voff, sn, ss;
  std::vector<cplx> W, V;
  int Mmax;
};

void build_fft(const HostGeom &hg, HostFFT &h) {
  const int np = hg.npair;
  h.M.resize(np); h.nphi = hg.nphi; h.shifted = hg.shifted; h.voff.assign(np, -1); h.sn.resize(np); h.ss.resize(np);
  h.order.resize(np);
  long long vtot = 0;
  int Mmax = 16;
  for (int ip = 0; ip < np; ++ip) {
    const int q = hg.nphi[ip] / 4;
    h.sn[ip] = hg.start_n[ip]; h.ss[ip] = hg.start_s[ip]; h.order[ip] = ip;
    if (q <= kTinyQ) h.M[ip] = 0;
    else if ((q & (q - 1)) == 0) h.M[ip] = q;
    else { h.M[ip] = nextpow2(2 * q - 1); h.voff[ip] = vtot; vtot += h.M[ip]; }
    Mmax = std::max(Mmax, h.M[ip]);
  }
  h.Mmax = Mmax;
  h.W.resize(Mmax);
  for (int k = 0; k < Mmax; ++k) {
    long double a = -2.0L * 3.14159265358979323846264338327950288L * k / Mmax;
    h.W[k] = mk((double)cosl(a), (double)sinl(a));
  }
  h.V.assign(std::max<long long>(vtot, 1), mk(0, 0));
  DevFFT &f = h.f;
  f.nside = hg.nside; f.npair = np; f.nring = hg.nring; f.Wn = Mmax; f.W = h.W.data(); f.V = h.V.data();
  f.voff = h.voff.data(); f.M = h.M.data(); f.nphi = h.nphi.data(); f.shifted = h.shifted.data();
  f.start_n = h.sn.data(); f.start_s = h.ss.data(); f.order = h.order.data(); f.mtop = nullptr; f.dist_n = 0; f.nb4_maxm = 1024; f.nb1_minm = 8192; f.pix.n = 0;
  std::vector<cplx> buf(Mmax + Mmax / 4);
  for (int ip = 0; ip < np; ++ip) bluestein_setup_body(BlockCtx(), f, ip, h.V.data(), buf.data());
}
}  // namespace

// nbatch = 2 or 4 (DFTs of one ring, or of both rings of the pair, per batch) -- both are exercised by the tests
extern "C" int emul_ring_synth(int nside, int mmax, int pitch, const cplx *X, double *map, int nbatch) {
  HostGeom hg = make_geom(nside);
  HostFFT h; build_fft(hg, h);
  std::vector<cplx> buf((size_t)nbatch * h.Mmax + h.Mmax / 4);
  for (int ip = 0; ip < hg.npair; ++ip) ring_synth_body(BlockCtx(), h.f, ip, X, pitch, mmax, map, buf.data(), nbatch);
  return 0;
}
extern "C" int emul_ring_anal(int nside, int mmax, int pitch, const double *map, cplx *X, int nbatch) {
  HostGeom hg = make_geom(nside);
  HostFFT h; build_fft(hg, h);
  std::vector<cplx> buf((size_t)nbatch * h.Mmax + h.Mmax / 4);
  const double w = 4.0 * M_PI / (double)hg.npix;
  for (int ip = 0; ip < hg.npair; ++ip) ring_anal_body(BlockCtx(), h.f, ip, map, X, pitch, mmax, w, buf.data(), nbatch);
  return 0;
}

namespace {
struct HostSpin {
  SpinTables t;
  std::vector<double2> uv;
  DevSpin d;
  DevGeom g;
};
void build_spin(const HostGeom &hg, int spin, int lmax, HostSpin &hs) {
  hs.t = make_spin_tables(spin, lmax, lmax);
  hs.uv.resize(hs.t.U.size());
  for (size_t i = 0; i < hs.uv.size(); ++i) hs.uv[i] = make_double2(hs.t.U[i], hs.t.V[i]);
  DevSpin &d = hs.d;
  d.spin = spin; d.lmax = lmax; d.mmax = lmax; d.thr_exp = kSeedThrExp; d.UV = hs.uv.data(); d.alpha = hs.t.alpha.data();
  d.k_hi = hs.t.k_hi.data(); d.k_lo = hs.t.k_lo.data(); d.k_e = hs.t.k_e.data(); d.pc = hs.t.pc.data();
  d.ps = hs.t.ps.data(); d.sg_p = hs.t.sg_p.data(); d.sg_m = hs.t.sg_m.data();
  DevGeom &g = hs.g;
  g.nside = hg.nside; g.npair = hg.npair; g.nring = hg.nring; g.npix = hg.npix; g.cth = hg.cth.data();
  g.sh_hi = hg.sh_hi.data(); g.sh_lo = hg.sh_lo.data(); g.ch_hi = hg.ch_hi.data(); g.ch_lo = hg.ch_lo.data();
}
}  // namespace

// seeds only: ks[m*npair+ip], s0..s3
extern "C" int emul_seeds(int nside, int lmax, int spin, int *ks, double *s0, double *s1, double *s2, double *s3) {
  HostGeom hg = make_geom(nside);
  HostSpin hs; build_spin(hg, spin, lmax, hs);
  for (int m = 0; m <= lmax; ++m)
    for (int ip = 0; ip < hg.npair; ++ip) {
      size_t o = (size_t)m * hg.npair + ip;
      if (spin) seed_one<true>(hs.g, hs.d, m, ip, ks[o], s0[o], s1[o], s2[o], s3[o]);
      else seed_one<false>(hs.g, hs.d, m, ip, ks[o], s0[o], s1[o], s2[o], s3[o]);
    }
  return 0;
}

extern "C" int emul_legendre_synth(int nside, int lmax, int spin, const cplx *a1, const cplx *a2, int pitch, cplx *X1,
                                   cplx *X2) {
  HostGeom hg = make_geom(nside);
  HostSpin hs; build_spin(hg, spin, lmax, hs);
  const double sgs = (spin & 1) ? -1.0 : 1.0;
  for (int m = 0; m <= lmax; ++m) {
    const int l0 = m > spin ? m : spin;
    const int K = lmax - l0 + 1;
    if (K <= 0) continue;
    const int64_t row = alm_idx(lmax, l0, m);
    for (int ip = 0; ip < hg.npair; ++ip) {
      int ks; double pm, pc, qm, qc;
      if (spin) seed_one<true>(hs.g, hs.d, m, ip, ks, pm, pc, qm, qc);
      else seed_one<false>(hs.g, hs.d, m, ip, ks, pm, pc, qm, qc);
      const double x = hg.cth[ip];
      double a0r = 0, a0i = 0, a1r = 0, a1i = 0, b0r = 0, b0i = 0, b1r = 0, b1i = 0;
      for (int k = ks; k < K; ++k) {
        const double al = hs.t.alpha[row + k];
        const double2 uv = hs.uv[row + k];
        const double sig = (k & 1) ? -1.0 : 1.0;
        if (!spin) {
          const cplx a = a1[row + k];
          if (k & 1) { a1r = fma(al * a.x, pc, a1r); a1i = fma(al * a.y, pc, a1i); }
          else { a0r = fma(al * a.x, pc, a0r); a0i = fma(al * a.y, pc, a0i); }
          double n = fma(x * uv.x, pc, -pm); pm = pc; pc = n;
        } else {
          const cplx g = a1[row + k], c = a2 ? a2[row + k] : mk(0, 0);
          const double gr = al * g.x, gi = al * g.y, cr = al * c.x, ci = al * c.y;
          const double hpr = -0.5 * (gr - ci), hpi = -0.5 * (gi + cr), hmr = -0.5 * sgs * (gr + ci), hmi = -0.5 * sgs * (gi - cr);
          a0r = fma(hpr, pc, a0r); a0i = fma(hpi, pc, a0i);
          b1r = fma(sig * hmr, pc, b1r); b1i = fma(sig * hmi, pc, b1i);
          a1r = fma(hmr, qc, a1r); a1i = fma(hmi, qc, a1i);
          b0r = fma(sig * hpr, qc, b0r); b0i = fma(sig * hpi, qc, b0i);
          double np = fma(fma(x, uv.x, uv.y), pc, -pm); pm = pc; pc = np;
          double nq = fma(fma(x, uv.x, -uv.y), qc, -qm); qm = qc; qc = nq;
        }
      }
      const double sg0 = ((l0 + m) & 1) ? -1.0 : 1.0;
      const int rn = ip, rs = hg.nring - 1 - ip;
      if (!spin) {
        X1[(size_t)rn * pitch + m] = mk(a0r + a1r, a0i + a1i);
        if (rs != rn) X1[(size_t)rs * pitch + m] = mk(sg0 * (a0r - a1r), sg0 * (a0i - a1i));
      } else {
        X1[(size_t)rn * pitch + m] = mk(a0r + a1r, a0i + a1i);
        X2[(size_t)rn * pitch + m] = mk(a0i - a1i, -(a0r - a1r));
        if (rs != rn) {
          X1[(size_t)rs * pitch + m] = mk(sg0 * (b0r + b1r), sg0 * (b0i + b1i));
          X2[(size_t)rs * pitch + m] = mk(sg0 * (b0i - b1i), -sg0 * (b0r - b1r));
        }
      }
    }
  }
  return 0;
}

extern "C" int emul_legendre_anal(int nside, int lmax, int spin, int pitch, const cplx *X1, const cplx *X2, cplx *o1,
                                  cplx *o2) {
  HostGeom hg = make_geom(nside);
  HostSpin hs; build_spin(hg, spin, lmax, hs);
  const double sgs = (spin & 1) ? -1.0 : 1.0;
  for (int m = 0; m <= lmax; ++m) {
    const int l0 = m > spin ? m : spin;
    const int K = lmax - l0 + 1;
    for (int l = m; l < l0 && l <= lmax; ++l) { o1[alm_idx(lmax, l, m)] = mk(0, 0); if (spin) o2[alm_idx(lmax, l, m)] = mk(0, 0); }
    if (K <= 0) continue;
    const int64_t row = alm_idx(lmax, l0, m);
    std::vector<double> S((size_t)K * 4, 0.0);
    const double sg0 = ((l0 + m) & 1) ? -1.0 : 1.0;
    for (int ip = 0; ip < hg.npair; ++ip) {
      int ks; double pm, pc, qm, qc;
      if (spin) seed_one<true>(hs.g, hs.d, m, ip, ks, pm, pc, qm, qc);
      else seed_one<false>(hs.g, hs.d, m, ip, ks, pm, pc, qm, qc);
      if (ks >= K) continue;
      const double x = hg.cth[ip];
      const int rn = ip, rs = hg.nring - 1 - ip;
      const cplx n1 = X1[(size_t)rn * pitch + m], s1 = rs != rn ? X1[(size_t)rs * pitch + m] : mk(0, 0);
      double f0r, f0i, f1r, f1i, f2r = 0, f2i = 0, f3r = 0, f3i = 0;
      if (!spin) {
        f0r = n1.x + sg0 * s1.x; f0i = n1.y + sg0 * s1.y; f1r = n1.x - sg0 * s1.x; f1i = n1.y - sg0 * s1.y;
      } else {
        const cplx n2 = X2[(size_t)rn * pitch + m], s2 = rs != rn ? X2[(size_t)rs * pitch + m] : mk(0, 0);
        f0r = n1.x - n2.y; f0i = n1.y + n2.x; f1r = n1.x + n2.y; f1i = n1.y - n2.x;
        f2r = sg0 * (s1.x - s2.y); f2i = sg0 * (s1.y + s2.x); f3r = sg0 * (s1.x + s2.y); f3i = sg0 * (s1.y - s2.x);
      }
      for (int k = ks; k < K; ++k) {
        const double2 uv = hs.uv[row + k];
        const double sig = (k & 1) ? -1.0 : 1.0;
        if (!spin) {
          if (k & 1) { S[k * 4 + 0] += pc * f1r; S[k * 4 + 1] += pc * f1i; }
          else { S[k * 4 + 0] += pc * f0r; S[k * 4 + 1] += pc * f0i; }
          double n = fma(x * uv.x, pc, -pm); pm = pc; pc = n;
        } else {
          S[k * 4 + 0] += pc * f0r + sig * qc * f2r; S[k * 4 + 1] += pc * f0i + sig * qc * f2i;
          S[k * 4 + 2] += qc * f1r + sig * pc * f3r; S[k * 4 + 3] += qc * f1i + sig * pc * f3i;
          double np = fma(fma(x, uv.x, uv.y), pc, -pm); pm = pc; pc = np;
          double nq = fma(fma(x, uv.x, -uv.y), qc, -qm); qm = qc; qc = nq;
        }
      }
    }
    for (int k = 0; k < K; ++k) {
      const double al = hs.t.alpha[row + k];
      if (!spin) o1[row + k] = mk(al * S[k * 4], al * S[k * 4 + 1]);
      else {
        const double pr = S[k * 4], pi = S[k * 4 + 1], mr = sgs * S[k * 4 + 2], mi = sgs * S[k * 4 + 3];
        o1[row + k] = mk(-0.5 * al * (pr + mr), -0.5 * al * (pi + mi));
        o2[row + k] = mk(-0.5 * al * (pi - mi), 0.5 * al * (pr - mr));
      }
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------- HEALPix degrade, Philox
#include "../../plancklens_b200/csrc/plk_blas.cuh"
#include "../../plancklens_b200/csrc/plk_rng.cuh"

// the index arithmetic of udgrade_sum_kernel, one output pixel after the other
extern "C" int emul_udgrade_sum(int nside_in, const double *in, int nside_out, double *out) {
  const long long npo = 12LL * nside_out * nside_out;
  const int f = nside_in / nside_out;
  for (long long po = 0; po < npo; ++po) {
    int ix, iy, face;
    hpx_ring2xyf(nside_out, po, ix, iy, face);
    out[po] = hpx_children_sum(in, nside_in, nside_out, ix, iy, face);
  }
  (void)f;
  return 0;
}
// ring -> (face, x, y) -> ring must be the identity
extern "C" long long emul_hpx_roundtrip_errors(int nside) {
  long long bad = 0;
  for (long long p = 0; p < 12LL * nside * nside; ++p) {
    int ix, iy, face;
    hpx_ring2xyf(nside, p, ix, iy, face);
    if (ix < 0 || iy < 0 || ix >= nside || iy >= nside || face < 0 || face > 11 || hpx_xyf2ring(nside, ix, iy, face) != p) ++bad;
  }
  return bad;
}
extern "C" int emul_philox_words(unsigned long long seed, unsigned long long stream, long long ncalls, unsigned int *out) {
  for (long long i = 0; i < ncalls; ++i) {
    const Philox4 p = philox4x32_10((uint64_t)i, stream, seed);
    for (int k = 0; k < 4; ++k) out[4 * i + k] = p.v[k];
  }
  return 0;
}
