/* CPU oracle: Legendre stage of the spin-weighted SHT on iso-latitude ring pairs.
 * TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product library.
 * PARITY UNPINNED at this seam (see oracle/__init__.py): the reference delegates the transform to
 * healpy / ducc0 (/root/reference/plancklens/shts.py:4-35), which are not vendored.
 *
 * This is a plain-C restatement of the published algorithm those libraries implement
 * (Reinecke & Seljebotn 2013, "libsharp"): per (m, ring pair) a three-term recurrence in l for
 *     slam_l(theta) = (-1)^m sqrt((2l+1)/4pi) d^l_{-m,s}(theta),
 *     E_{l+1} lam_{l+1} = (x - mu_l) lam_l - E_l lam_{l-1},   mu_l = m1 m2 / (l(l+1)),
 *     E_l = sqrt((l^2-m1^2)(l^2-m2^2) / (l^2 (4 l^2 - 1))),   (m1,m2) = (-m, +-s),
 * started at l0 = max(m,|s|) from the closed form of d^{l0}, with an explicit power-of-two scale
 * while the value is below the double range, and the north/south identity
 *     slam_lm(pi-theta) = (-1)^{l+m}  (-s)lam_lm(theta).
 * Deliberately different in normalisation and start-up from the CUDA kernels it checks.
 *
 * Conventions (reference plancklens/utils_spin.py:1-14):
 *   +s a_lm = -(G + iC),  -s a_lm = -(-1)^s (G - iC);   maps = Re, Im of sum +s a_lm +sY_lm.
 * Phase arrays are [ring][m] complex (row pitch mmax+1):
 *   synthesis output  X1_m, X2_m with  map(phi) = X_0 + 2 Re sum_{m>0} X_m e^{i m phi}
 *   analysis  input   X1_m, X2_m = sum_j map_j e^{-i m phi_j}  (weights applied here).
 */
#define _GNU_SOURCE
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex dcmplx;

#define SCALE_STEP 200
#define START_EXP (-900)

typedef struct {
  int l0;
  double *E;     /* E[l], l = 0..lmax+1 */
  double *invE;  /* 1/E[l] (0 where E==0) */
  double *mu;    /* |m1 m2| / (l(l+1)) */
  long double lognorm; /* natural log of the seed normalisation incl. sqrt((2l0+1)/4pi) */
  int pc, ps;    /* powers of cos(theta/2), sin(theta/2) in d^{l0}_{-m,+s} ; swapped for -s */
  double sgn_p, sgn_m; /* signs of the +s and -s seeds (including (-1)^m) */
} mtab;

static void mtab_init(mtab *t, int lmax, int m, int s) {
  int l0 = m > s ? m : s;
  t->l0 = l0;
  t->E = (double *)calloc(lmax + 3, sizeof(double));
  t->invE = (double *)calloc(lmax + 3, sizeof(double));
  t->mu = (double *)calloc(lmax + 3, sizeof(double));
  for (int l = 1; l <= lmax + 1; ++l) {
    long double ll = (long double)l * l;
    long double num = (ll - (long double)m * m) * (ll - (long double)s * s);
    long double den = ll * (4.0L * ll - 1.0L);
    long double e = num > 0 ? sqrtl(num / den) : 0.0L;
    t->E[l] = (double)e;
    t->invE[l] = e > 0 ? (double)(1.0L / e) : 0.0;
    t->mu[l] = (double)((long double)m * s / ((long double)l * (l + 1)));
  }
  /* seed: d^{l0}_{-m,+s}.  N(j,k) = sqrt((2j)!/((j+k)!(j-k)!)) */
  int j = l0, k = (m >= s) ? s : m;
  t->lognorm = 0.5L * (lgammal(2.0L * j + 1) - lgammal((long double)j + k + 1) - lgammal((long double)j - k + 1))
             + 0.5L * logl((2.0L * j + 1) / (4.0L * M_PIl));
  double sm = (m & 1) ? -1.0 : 1.0;
  if (m >= s) {
    /* d^j_{-j,m2} = N(j,m2) cos^{j-m2} sin^{j+m2}, j = m */
    t->pc = m - s; t->ps = m + s;       /* for m2=+s ; for m2=-s the two swap */
    t->sgn_p = sm; t->sgn_m = sm;
  } else {
    /* j = s > m, m1 = -m:
       m2=+j: d^j_{m1,j}  = N(j,m1) cos^{j+m1} sin^{j-m1}        = N cos^{s-m} sin^{s+m}
       m2=-j: d^j_{m1,-j} = N(j,m1) cos^{j-m1} (-sin)^{j+m1}     = N cos^{s+m} (-1)^{s-m} sin^{s-m} */
    t->pc = s - m; t->ps = s + m;
    t->sgn_p = sm; t->sgn_m = sm * (((s - m) & 1) ? -1.0 : 1.0);
  }
}

static void mtab_free(mtab *t) { free(t->E); free(t->invE); free(t->mu); }

/* Runs the recurrence for one sign (musgn = +1 for m2=+s since -m1 m2 = +m s, -1 for m2=-s) from l0 until the
 * true value is representable; returns the first l (>= l0) from which plain doubles can be used, with
 * lam[l-1] in *pm and lam[l] in *pc.  Returns lmax+1 if never reached. */
static int ramp(const mtab *t, int lmax, double x, double musgn, long double logseed, double sgn,
                double *pm, double *pc) {
  /* seed = sgn * exp(logseed) = sgn * 2^(e0) * f */
  long double l2 = logseed / M_LN2l;
  long e = (long)floorl(l2);
  double f = sgn * (double)exp2l(l2 - (long double)e);
  int l = t->l0;
  double vm = 0.0, vc = f;
  while (e < START_EXP) {
    if (l >= lmax) return lmax + 1;
    double vn = ((x + musgn * t->mu[l]) * vc - t->E[l] * vm) * t->invE[l + 1];
    vm = vc; vc = vn; ++l;
    if (fabs(vc) > ldexp(1.0, SCALE_STEP)) {
      vm = ldexp(vm, -SCALE_STEP); vc = ldexp(vc, -SCALE_STEP); e += SCALE_STEP;
    }
  }
  *pm = ldexp(vm, (int)e);
  *pc = ldexp(vc, (int)e);
  return l;
}

static inline long double logseed_of(const mtab *t, long double logch, long double logsh, int pc, int ps) {
  long double r = t->lognorm;
  if (pc) r += (long double)pc * logch;
  if (ps) r += (long double)ps * logsh;
  return r;
}

/* ring pair ip: north ring index rn[ip] (row in phase arrays), south rs[ip] (or -1), x = cos(theta_north),
 * chalf/shalf = natural log of cos/sin(theta_north/2) in long double (so that sin^m is exact to ~m*1e-19).  mstep > 1 computes only every mstep-th m (bounded CPU-baseline sample). */
/* Common body: m runs over mlist[0..nm) when mlist != NULL, else over 0, mstep, 2 mstep, ... <= mmax. */
static int synth_body(int spin, int lmax, int mmax, int npair, const int *rn, const int *rs,
                      const double *cth, const long double *chalf, const long double *shalf,
                      const dcmplx *almG, const dcmplx *almC, dcmplx *X1, dcmplx *X2, int mstep,
                      const int *mlist, int nm) {
  if (!mlist) nm = mmax / mstep + 1;
  const int pitch = mlist ? nm : mmax + 1;   /* m-list form: compact phase arrays [ring][nm], column = position in mlist */
#pragma omp parallel for schedule(dynamic, 1)
  for (int im = 0; im < nm; ++im) {
    const int m = mlist ? mlist[im] : im * mstep;
    const int col = mlist ? im : m;
    mtab t; mtab_init(&t, lmax, m, spin);
    const dcmplx *g = almG + (size_t)m * (2 * lmax + 1 - m) / 2;   /* g[l] valid for l>=m */
    const dcmplx *c = almC ? almC + (size_t)m * (2 * lmax + 1 - m) / 2 : NULL;
    for (int ip = 0; ip < npair; ++ip) {
      double x = cth[ip];
      if (spin == 0) {
        double pm, pc;
        int l = ramp(&t, lmax, x, 0.0, logseed_of(&t, chalf[ip], shalf[ip], t.pc, t.ps), t.sgn_p, &pm, &pc);
        dcmplx ev = 0, od = 0;
        for (; l <= lmax; ++l) {
          if ((l - m) & 1) od += g[l] * pc; else ev += g[l] * pc;
          double pn = (x * pc - t.E[l] * pm) * t.invE[l + 1];
          pm = pc; pc = pn;
        }
        X1[(size_t)rn[ip] * pitch + col] = ev + od;
        if (rs[ip] >= 0) X1[(size_t)rs[ip] * pitch + col] = ev - od;
      } else {
        double ppm, ppc, mpm, mpc;
        int lp = ramp(&t, lmax, x, +1.0, logseed_of(&t, chalf[ip], shalf[ip], t.pc, t.ps), t.sgn_p, &ppm, &ppc);
        int lm = ramp(&t, lmax, x, -1.0, logseed_of(&t, chalf[ip], shalf[ip], t.ps, t.pc), t.sgn_m, &mpm, &mpc);
        double ssgn = (spin & 1) ? -1.0 : 1.0;
        dcmplx An_p = 0, An_m = 0, As_p = 0, As_m = 0;   /* A^{+-} on north / south ring */
        int l = lp < lm ? lp : lm;
        for (; l <= lmax; ++l) {
          dcmplx ap = -(g[l] + I * c[l]);
          dcmplx am = -ssgn * (g[l] - I * c[l]);
          double sig = ((l + m) & 1) ? -1.0 : 1.0;
          if (l >= lp) {
            An_p += ap * ppc; As_m += sig * am * ppc;
            double pn = ((x + t.mu[l]) * ppc - t.E[l] * ppm) * t.invE[l + 1];
            ppm = ppc; ppc = pn;
          }
          if (l >= lm) {
            An_m += am * mpc; As_p += sig * ap * mpc;
            double pn = ((x - t.mu[l]) * mpc - t.E[l] * mpm) * t.invE[l + 1];
            mpm = mpc; mpc = pn;
          }
        }
        size_t in = (size_t)rn[ip] * pitch + col;
        X1[in] = 0.5 * (An_p + An_m);
        X2[in] = -0.5 * I * (An_p - An_m);
        if (rs[ip] >= 0) {
          size_t is = (size_t)rs[ip] * pitch + col;
          X1[is] = 0.5 * (As_p + As_m);
          X2[is] = -0.5 * I * (As_p - As_m);
        }
      }
    }
    mtab_free(&t);
  }
  return 0;
}

int csht_synth(int spin, int lmax, int mmax, int npair, const int *rn, const int *rs,
               const double *cth, const long double *chalf, const long double *shalf,
               const dcmplx *almG, const dcmplx *almC, dcmplx *X1, dcmplx *X2, int mstep) {
  return synth_body(spin, lmax, mmax, npair, rn, rs, cth, chalf, shalf, almG, almC, X1, X2, mstep, NULL, 0);
}
/* only the m of mlist[0..nm), with COMPACT phase arrays X[ring][nm] (column = position in mlist): full-size parity
 * samples that must include m ~ lmax and chosen m in between */
int csht_synth_mlist(int spin, int lmax, int mmax, int npair, const int *rn, const int *rs,
                     const double *cth, const long double *chalf, const long double *shalf,
                     const dcmplx *almG, const dcmplx *almC, dcmplx *X1, dcmplx *X2, const int *mlist, int nm) {
  return synth_body(spin, lmax, mmax, npair, rn, rs, cth, chalf, shalf, almG, almC, X1, X2, 1, mlist, nm);
}

/* weight[ip] multiplies both rings of the pair (4 pi / npix for HEALPix). */
static int anal_body(int spin, int lmax, int mmax, int npair, const int *rn, const int *rs,
                     const double *cth, const long double *chalf, const long double *shalf, const double *weight,
                     const dcmplx *X1, const dcmplx *X2, dcmplx *almG, dcmplx *almC, int mstep,
                     const int *mlist, int nm) {
  if (!mlist) nm = mmax / mstep + 1;
  const int pitch = mlist ? nm : mmax + 1;   /* m-list form: compact phase arrays [ring][nm], column = position in mlist */
#pragma omp parallel for schedule(dynamic, 1)
  for (int im = 0; im < nm; ++im) {
    const int m = mlist ? mlist[im] : im * mstep;
    const int col = mlist ? im : m;
    mtab t; mtab_init(&t, lmax, m, spin);
    dcmplx *g = almG + (size_t)m * (2 * lmax + 1 - m) / 2;
    dcmplx *c = almC ? almC + (size_t)m * (2 * lmax + 1 - m) / 2 : NULL;
    dcmplx *ap = (dcmplx *)calloc(lmax + 1, sizeof(dcmplx));
    dcmplx *am = (dcmplx *)calloc(lmax + 1, sizeof(dcmplx));
    for (int ip = 0; ip < npair; ++ip) {
      double x = cth[ip], w = weight[ip];
      size_t in = (size_t)rn[ip] * pitch + col;
      if (spin == 0) {
        dcmplx fn = w * X1[in], fs = rs[ip] >= 0 ? w * X1[(size_t)rs[ip] * pitch + col] : 0;
        dcmplx fe = fn + fs, fo = fn - fs;
        double pm, pc;
        int l = ramp(&t, lmax, x, 0.0, logseed_of(&t, chalf[ip], shalf[ip], t.pc, t.ps), t.sgn_p, &pm, &pc);
        for (; l <= lmax; ++l) {
          ap[l] += pc * (((l - m) & 1) ? fo : fe);
          double pn = (x * pc - t.E[l] * pm) * t.invE[l + 1];
          pm = pc; pc = pn;
        }
      } else {
        dcmplx pn_ = w * (X1[in] + I * X2[in]), mn_ = w * (X1[in] - I * X2[in]);
        dcmplx ps_ = 0, ms_ = 0;
        if (rs[ip] >= 0) {
          size_t is = (size_t)rs[ip] * pitch + col;
          ps_ = w * (X1[is] + I * X2[is]); ms_ = w * (X1[is] - I * X2[is]);
        }
        double ppm, ppc, mpm, mpc;
        int lp = ramp(&t, lmax, x, +1.0, logseed_of(&t, chalf[ip], shalf[ip], t.pc, t.ps), t.sgn_p, &ppm, &ppc);
        int lm = ramp(&t, lmax, x, -1.0, logseed_of(&t, chalf[ip], shalf[ip], t.ps, t.pc), t.sgn_m, &mpm, &mpc);
        int l = lp < lm ? lp : lm;
        for (; l <= lmax; ++l) {
          double sig = ((l + m) & 1) ? -1.0 : 1.0;
          if (l >= lp) {
            ap[l] += ppc * pn_; am[l] += sig * ppc * ms_;
            double pn = ((x + t.mu[l]) * ppc - t.E[l] * ppm) * t.invE[l + 1];
            ppm = ppc; ppc = pn;
          }
          if (l >= lm) {
            am[l] += mpc * mn_; ap[l] += sig * mpc * ps_;
            double pn = ((x - t.mu[l]) * mpc - t.E[l] * mpm) * t.invE[l + 1];
            mpm = mpc; mpc = pn;
          }
        }
      }
    }
    if (spin == 0) {
      for (int l = m; l <= lmax; ++l) g[l] = ap[l];
    } else {
      double ssgn = (spin & 1) ? -1.0 : 1.0;
      for (int l = m; l <= lmax; ++l) {
        dcmplx a = ap[l], b = ssgn * am[l];
        g[l] = -0.5 * (a + b);
        c[l] = 0.5 * I * (a - b);
      }
    }
    free(ap); free(am);
    mtab_free(&t);
  }
  return 0;
}

int csht_anal(int spin, int lmax, int mmax, int npair, const int *rn, const int *rs,
              const double *cth, const long double *chalf, const long double *shalf, const double *weight,
              const dcmplx *X1, const dcmplx *X2, dcmplx *almG, dcmplx *almC, int mstep) {
  return anal_body(spin, lmax, mmax, npair, rn, rs, cth, chalf, shalf, weight, X1, X2, almG, almC, mstep, NULL, 0);
}
int csht_anal_mlist(int spin, int lmax, int mmax, int npair, const int *rn, const int *rs,
                    const double *cth, const long double *chalf, const long double *shalf, const double *weight,
                    const dcmplx *X1, const dcmplx *X2, dcmplx *almG, dcmplx *almC, const int *mlist, int nm) {
  return anal_body(spin, lmax, mmax, npair, rn, rs, cth, chalf, shalf, weight, X1, X2, almG, almC, 1, mlist, nm);
}

int csht_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
