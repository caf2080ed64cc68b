// Host-side table generation for a (nside, lmax, mmax) plan: ring-pair geometry and, per spin, the
// normalised three-term recurrence coefficients.  All arithmetic in x87 long double, rounded once.
//
// Recurrence (general spin, (m1,m2) = (-m, +-s)), see DESIGN.md section "Legendre recurrence":
//   slam_l(theta) = (-1)^m sqrt((2l+1)/4pi) d^l_{-m,s}(theta)
//   E_{l+1} lam_{l+1} = (x - m1 m2/(l(l+1))) lam_l - E_l lam_{l-1},  E_l^2 = (l^2-m^2)(l^2-s^2)/(l^2(4l^2-1))
// With p_l = lam_l / alpha_l and alpha_{l+1} = alpha_{l-1} E_l / E_{l+1} the coefficient of p_{l-1} is exactly 1:
//   p_{l+1} = (x U_l +- V_l) p_l - p_{l-1},   U_l = alpha_l / (E_{l+1} alpha_{l+1}),  V_l = m s/(l(l+1)) U_l
// (+V for the +s function, -V for the -s one).  Two FMAs per step instead of three operations.
#pragma once
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "plk_common.h"

namespace plk {

struct HostGeom {
  int nside, npair, nring;
  int64_t npix;
  std::vector<double> cth;             // cos(theta) of the north ring of the pair
  std::vector<double> sh_hi, sh_lo;    // sin(theta/2) as double-double
  std::vector<double> ch_hi, ch_lo;    // cos(theta/2)
  std::vector<double> sth;             // sin(theta) (for mlim estimates only)
  std::vector<int> nphi;               // pixels per ring
  std::vector<int64_t> start_n, start_s;  // first pixel of north / south ring (south = -1 on the equator)
  std::vector<int> shifted;            // 1: phi0 = pi/nphi, 0: phi0 = 0
};

inline void split_ld(long double v, double &hi, double &lo) {
  hi = (double)v;
  lo = (double)(v - (long double)hi);
}

inline HostGeom make_geom(int nside) {
  HostGeom g;
  const int N = nside;
  g.nside = N; g.npair = 2 * N; g.nring = 4 * N - 1; g.npix = 12LL * N * N;
  g.cth.resize(g.npair); g.sh_hi.resize(g.npair); g.sh_lo.resize(g.npair);
  g.ch_hi.resize(g.npair); g.ch_lo.resize(g.npair); g.sth.resize(g.npair);
  g.nphi.resize(g.npair); g.start_n.resize(g.npair); g.start_s.resize(g.npair); g.shifted.resize(g.npair);
  const long double NL = N;
  for (int ip = 0; ip < g.npair; ++ip) {
    const int i = ip + 1;  // ring number 1..2N (north cap, then belt down to the equator)
    long double omz;       // 1 - cos(theta)
    if (i < N) {
      omz = (long double)i * i / (3.0L * NL * NL);
      g.nphi[ip] = 4 * i;
      g.start_n[ip] = 2LL * i * (i - 1);
      g.start_s[ip] = g.npix - 2LL * i * (i + 1);
      g.shifted[ip] = 1;
    } else {
      omz = 1.0L - 2.0L * (2.0L * NL - i) / (3.0L * NL);
      g.nphi[ip] = 4 * N;
      g.start_n[ip] = 2LL * N * (N - 1) + (int64_t)(i - N) * 4 * N;
      const int is = 4 * N - i;  // southern twin
      g.start_s[ip] = (i == 2 * N) ? -1 : 2LL * N * (N - 1) + (int64_t)(is - N) * 4 * N;
      g.shifted[ip] = ((i - N) % 2 == 0) ? 1 : 0;
    }
    g.cth[ip] = (double)(1.0L - omz);
    split_ld(sqrtl(omz / 2.0L), g.sh_hi[ip], g.sh_lo[ip]);
    split_ld(sqrtl(1.0L - omz / 2.0L), g.ch_hi[ip], g.ch_lo[ip]);
    g.sth[ip] = (double)sqrtl(omz * (2.0L - omz));
  }
  return g;
}

struct SpinTables {
  int spin, lmax, mmax;
  std::vector<double> U, V;     // [alm_idx(l,m)], valid for l >= l0(m) = max(m,spin)
  std::vector<double> alpha;    // lam = alpha * p ; zero for l < l0
  // seed normalisation kappa(m) as (hi, lo, e): |lam^{+-}_{l0}| = kappa * ch^{pc} sh^{ps}  (powers swap for '-')
  std::vector<double> k_hi, k_lo;
  std::vector<int> k_e;
  std::vector<int> pc, ps;      // powers for the '+' function
  std::vector<signed char> sg_p, sg_m;  // signs of the '+' and '-' seeds
};

inline void fill_m(SpinTables &t, int m) {
  const int s = t.spin, lmax = t.lmax;
  const int l0 = m > s ? m : s;
  if (l0 > lmax) return;
  auto Efun = [&](int l) -> long double {
    if (l <= l0) return 0.0L;
    long double ll = (long double)l * l;
    long double num = (ll - (long double)m * m) * (ll - (long double)s * s);
    return sqrtl(num / (ll * (4.0L * ll - 1.0L)));
  };
  const int64_t base = alm_idx(lmax, 0, m);
  long double a_prev = 1.0L;   // alpha_{l-1}
  long double a_cur = 1.0L;    // alpha_l
  // alpha_{l0} = alpha_{l0+1} = 1 ; alpha_{l+1} = alpha_{l-1} E_l / E_{l+1} for l >= l0+1
  for (int l = l0; l <= lmax; ++l) {
    long double a_next;
    long double E1 = Efun(l + 1);
    if (l == l0) a_next = 1.0L;
    else a_next = a_prev * Efun(l) / E1;
    long double U = a_cur / (E1 * a_next);
    long double mu = (l > 0) ? (long double)m * s / ((long double)l * (l + 1)) : 0.0L;
    t.U[base + l] = (double)U;
    t.V[base + l] = (double)(mu * U);
    t.alpha[base + l] = (double)a_cur;
    a_prev = a_cur; a_cur = a_next;
  }
}

inline SpinTables make_spin_tables(int spin, int lmax, int mmax, int nthreads = 8) {
  SpinTables t;
  t.spin = spin; t.lmax = lmax; t.mmax = mmax;
  const int64_t n = alm_size(lmax, mmax);
  t.U.assign(n, 0.0); t.V.assign(n, 0.0); t.alpha.assign(n, 0.0);
  t.k_hi.assign(mmax + 1, 0.0); t.k_lo.assign(mmax + 1, 0.0); t.k_e.assign(mmax + 1, 0);
  t.pc.assign(mmax + 1, 0); t.ps.assign(mmax + 1, 0);
  t.sg_p.assign(mmax + 1, 1); t.sg_m.assign(mmax + 1, 1);
  std::vector<std::thread> th;
  for (int w = 0; w < nthreads; ++w)
    th.emplace_back([&t, w, nthreads, mmax]() {
      for (int m = w; m <= mmax; m += nthreads) fill_m(t, m);
    });
  for (auto &x : th) x.join();
  // seed constants: N(j,k) = sqrt((2j)!/((j+k)!(j-k)!)) = sqrt(C(2j, j-k)), j = max(m,s), k = min(m,s)
  // built by the ratio recurrence in m (m >= s):  C(2m,m-s)/C(2m-2,m-1-s) = 2m(2m-1)/((m+s)(m-s)).
  const int s = spin;
  long double Nms = 1.0L;  // N(m,s) for m >= s, started at m = s
  for (int m = 0; m <= mmax; ++m) {
    long double Nv;
    if (m < s) {
      // C(2s, s-m)
      long double c = 1.0L;
      for (int k = 1; k <= s - m; ++k) c = c * (long double)(2 * s - (s - m) + k) / (long double)k;
      Nv = sqrtl(c);
    } else {
      if (m > s) Nms *= sqrtl((long double)(2 * m) * (2 * m - 1) / ((long double)(m + s) * (m - s)));
      Nv = Nms;
    }
    const int l0 = m > s ? m : s;
    long double kap = Nv * sqrtl((2.0L * l0 + 1.0L) / (4.0L * 3.14159265358979323846264338327950288L));
    int e;
    long double f = frexpl(kap, &e);
    split_ld(f, t.k_hi[m], t.k_lo[m]);
    t.k_e[m] = e;
    const signed char sm = (m & 1) ? -1 : 1;
    if (m >= s) {
      // d^m_{-m,+s} = N cos^{m-s} sin^{m+s} ; d^m_{-m,-s} = N cos^{m+s} sin^{m-s}
      t.pc[m] = m - s; t.ps[m] = m + s; t.sg_p[m] = sm; t.sg_m[m] = sm;
    } else {
      // j = s: d^s_{-m,+s} = N cos^{s-m} sin^{s+m} ; d^s_{-m,-s} = N cos^{s+m} (-1)^{s-m} sin^{s-m}
      t.pc[m] = s - m; t.ps[m] = s + m; t.sg_p[m] = sm;
      t.sg_m[m] = (signed char)(sm * (((s - m) & 1) ? -1 : 1));
    }
  }
  return t;
}

}  // namespace plk
