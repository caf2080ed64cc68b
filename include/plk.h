/* libplk_b200: C ABI of the B200 (sm_100a) spin-weighted spherical-harmonic-transform hot path of
 * plancklens.  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * Every entry point returns 0 on success and a negative PLK_E* code otherwise (never throws);
 * plk_last_error() returns a human-readable message for the calling thread.
 *
 * alm layout: healpy m-major triangular order, complex128 (re, im interleaved), mmax == lmax:
 *   idx(l, m) = m (2 lmax + 1 - m) / 2 + l,  only m >= 0 stored.
 * map layout: HEALPix RING order, float64, npix = 12 nside^2.
 * Conventions (reference plancklens/utils_spin.py:1-14): spin 0 is the healpy scalar transform
 * T = sum a_lm Y_lm;  spin s > 0 uses  (+s)a_lm = -(G_lm + i C_lm)  and returns (Re, Im) of the spin-s field
 * (spin 2: (G, C) = (E, B) -> (Q, U)).  Analysis is the single-pass adjoint times 4 pi / npix
 * (healpy map2alm(iter=0), uniform weights) -- the only form the reference hot path calls.
 *
 * "_dev" functions take DEVICE pointers and a cudaStream_t (passed as void*); the caller owns all buffers.
 * "_host" functions take HOST pointers, stage through pinned memory and synchronise before returning.
 *
 * Reference interface each entry point replaces (paths relative to /root/reference):
 *   plk_alm2map_*   spin 0: plancklens/shts.py:12,35 (hp.alm2map; call sites qcinv/opfilt_tt.py:187, qest.py:471,514)
 *                   spin s: plancklens/shts.py:22,35 and utils_spin.py:21-27 (hp.alm2map_spin; opfilt_pp.py:260,
 *                           qest.py:464,504,530,593,636, utils_qe.py:70)
 *   plk_map2alm_*   spin 0: plancklens/shts.py:16,35 (hp.map2alm(iter=0); opfilt_tt.py:34,189, filt_simple.py:399)
 *                   spin s: plancklens/shts.py:26,35 and utils_spin.py:29-34 (hp.map2alm_spin; opfilt_pp.py:265,314,
 *                           qest.py:259,280, filt_simple.py:404, utils_qe.py:125)
 *   fl_* arguments  fuse the hp.almxfl calls that bracket those transforms (opfilt_tt.py:185,190, qest.py:260-262,
 *                   463, 494-503).
 *   plk_almxfl_dev, plk_alm_axpy_dev, plk_alm_dot_dev, plk_alm_copy_dev, plk_alm_splice_dev
 *                   hp.almxfl; cd_solve.py:75-86 vector updates; opfilt_tt.py:43-51 / opfilt_pp.py:27-34 dot_op;
 *                   qcinv/util_alm.py:8-44 alm_copy / alm_splice.
 *   plk_map_*       numpy per-pixel passes: opfilt_tt.py:193-205 (N^-1 and template projection),
 *                   opfilt_pp.py:272-303, qest.py:256-257, 276-278 (QE leg products).
 */
#ifndef PLK_H
#define PLK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct plk_plan plk_plan;

#define PLK_OK 0
#define PLK_EINVAL (-1)   /* bad argument */
#define PLK_ECUDA (-2)    /* CUDA runtime error (see plk_last_error) */
#define PLK_ENOMEM (-3)
#define PLK_ENODEV (-4)   /* no CUDA device / not sm_100 */

const char *plk_last_error(void);
int plk_version(void);
/* number of CUDA kernels this library has launched in the calling process (all plans) */
long long plk_launch_count(void);

/* Builds geometry, ring-FFT tables and scratch for (nside, lmax, mmax) on the CURRENT device.
 * nside: power of two, 1 <= nside <= 8192.  mmax must equal lmax (the only case the reference uses). */
int plk_plan_create(plk_plan **plan, int nside, int lmax, int mmax);
int plk_plan_destroy(plk_plan *plan);
/* bytes of device memory currently held by the plan (tables + scratch) */
long long plk_plan_device_bytes(const plk_plan *plan);
/* Start threshold 2^exp2 of the on-the-fly Legendre recurrences (default 2^-60, the value at which libsharp, which
 * healpy / ducc0 wrap behind plancklens/shts.py:33-35, starts accumulating).  (l, m, ring) contributions below it are skipped
 * near the poles; results must not depend on it at the 1e-10 level, which the full-size parity tests check by
 * varying it.  Rebuilds the per-spin seed tables on next use; exp2 in [-900, -20]. */
int plk_plan_set_seed_threshold(plk_plan *plan, int exp2);
int plk_plan_nside(const plk_plan *plan);
int plk_plan_lmax(const plk_plan *plan);

/* ---- transforms, device pointers.  alm2/map2/fl2 are ignored for spin 0.  alm2 may be NULL for spin > 0
 *      (zero curl component).  fl1/fl2: optional per-l real factors (device, lmax+1 doubles) multiplied onto
 *      alm1/alm2 before synthesis (after analysis); NULL = 1. */
int plk_alm2map_dev(plk_plan *plan, int spin, const void *alm1, const void *alm2, const double *fl1,
                    const double *fl2, double *map1, double *map2, void *stream);
int plk_map2alm_dev(plk_plan *plan, int spin, const double *map1, const double *map2, const double *fl1,
                    const double *fl2, void *alm1, void *alm2, void *stream);

/* Analysis with an additive per-l term folded into the output pass (no separate combine kernel):
 *   alm_c = fl_c[l] * analysis_c + afl_c[l] * add_c[l, m]      (afl_c: lmax + 1 doubles; add_c must not alias alm_c)
 * -- the forward operator of the CG filters, A x = S^-1 x + B^t N^-1 B x (qcinv/opfilt_tt.py:67-73, opfilt_pp.py:51-55
 * with a diagonal S^-1): add = x, afl = 1 / C_l.  add2 / afl2 are ignored for spin 0. */
int plk_map2alm_add_dev(plk_plan *plan, int spin, const double *map1, const double *map2, const double *fl1,
                        const double *fl2, const void *add1, const double *afl1, const void *add2, const double *afl2,
                        void *alm1, void *alm2, void *stream);

/* Analysis of maps that are never written to memory: pixel p of input component c is evaluated inside the ring kernel as
 *   sum_{k < nterm} scale[k] * a[k][p] * (b[k] ? b[k][p] : 1)          (device pointers, 32-byte aligned, npix doubles)
 * -- the per-pixel leg products of the quadratic estimators (qest.py:256-257: G t, C t; qest.py:276-278:
 * (Q - iU)(G3 + iC3) - (Q + iU)(G1 - iC1)) and the N^-1 multiply of the one-map polarization filter
 * (opfilt_pp.py:272-303) fused into the transform that consumes them.  add1 == NULL: no additive term. */
#define PLK_MAX_PIX_TERMS 6
/* Lanes: independent chains of launches issued concurrently from different host threads / streams (the temperature and
 * the polarization filter of one simulation; no counterpart in the reference, where filt_cinv.py:196-203 and :275-289 run
 * one after the other).  The lane is a thread-local index; each lane owns the library's reduction scratch it uses. */
#define PLK_MAX_LANES 4
int plk_set_lane(int lane);
int plk_get_lane(void);
typedef struct plk_pixprog {
  int nterm;
  const double *a[PLK_MAX_PIX_TERMS];
  const double *b[PLK_MAX_PIX_TERMS];
  double scale[PLK_MAX_PIX_TERMS];
} plk_pixprog;
int plk_map2alm_pix_dev(plk_plan *plan, int spin, const plk_pixprog *pix1, const plk_pixprog *pix2, const double *fl1,
                        const double *fl2, const void *add1, const double *afl1, const void *add2, const double *afl2,
                        void *alm1, void *alm2, void *stream);

/* ---- transforms, host pointers (numpy arrays on the Python side) */
int plk_alm2map_host(plk_plan *plan, int spin, const void *alm1, const void *alm2, double *map1, double *map2);
int plk_map2alm_host(plk_plan *plan, int spin, const double *map1, const double *map2, void *alm1, void *alm2);

/* ---- Legendre / ring-FFT stages on their own (profiling and tests).  X: [4 nside - 1][mmax + 1] complex128 */
int plk_legendre_synth_dev(plk_plan *plan, int spin, const void *alm1, const void *alm2, const double *fl1,
                           const double *fl2, void *X1, void *X2, void *stream);
int plk_legendre_anal_dev(plk_plan *plan, int spin, const void *X1, const void *X2, const double *fl1,
                          const double *fl2, void *alm1, void *alm2, void *stream);
int plk_ring_synth_dev(plk_plan *plan, const void *X, double *map, void *stream);
int plk_ring_anal_dev(plk_plan *plan, const double *map, void *X, void *stream);

/* ---- alm BLAS-1 (device pointers; n = number of complex coefficients of an lmax/mmax=lmax triangle) */
/* out[l,m] = fl[l] * in[l,m] (fl has nfl entries; l >= nfl multiplies by 0, as hp.almxfl); in == out allowed */
int plk_almxfl_dev(int lmax, const void *in, const double *fl, int nfl, void *out, void *stream);
/* y += a * x  (a read from device memory *a_dev if a_dev != NULL, else the host value a) */
int plk_alm_axpy_dev(long long n, double a, const double *a_dev, const void *x, void *y, void *stream);
/* result_dev[0] = sum_{l>=lmin} [ a_l0 b_l0 + 2 sum_{m>0} Re(a_lm conj b_lm) ]  (opfilt_tt/pp dot_op) */
int plk_alm_dot_dev(int lmax, int lmin, const void *a, const void *b, double *result_dev, void *stream);
/* two-component form (opfilt_pp.py:27-34: E and B summed), one device scalar */
int plk_alm_dot2_dev(int lmax, int lmin, const void *a1, const void *b1, const void *a2, const void *b2,
                     double *result_dev, void *stream);
/* n <= 4 components (opfilt_tp.py:46-58: T, E and B summed, all from lmin), host arrays of device pointers */
int plk_alm_dotn_dev(int lmax, int lmin, int n, const void *const *a, const void *const *b, double *result_dev,
                     void *stream);
/* One-kernel form of the dot products above (n <= 4 components) with the CG step-length arithmetic of
 * cd_solve.py:69-71, :95-99 folded into the final sum (the block that finishes last adds the per-m partials in a fixed
 * order, so results are reproducible):
 *   out3[0] = s = sum_j dot(a_j, b_j)
 *   num != NULL: out3[1] = scale * num[0] / s, out3[2] = -out3[1]     (alpha = (d.r)/(d.Ad) and -alpha)
 *   den != NULL: out3[1] = scale * s / den[0], out3[2] = -out3[1]     (beta = -(d'.Ad)/(d.Ad) with scale = -1)
 * num, den, out3 are device pointers; a zero divisor gives 0.  At most one of num / den. */
int plk_alm_dot_fused_dev(int lmax, int lmin, int n, const void *const *a, const void *const *b, const double *num,
                          const double *den, double scale, double *out3, void *stream);
/* y1 += a x1 ; y2 -= a x2 with a read from device memory: solution and residual update of one CG iteration
 * (cd_solve.py:75, :82-84) in one pass */
int plk_alm_axpy2_dev(long long n, const double *a_dev, const void *x1, void *y1, const void *x2, void *y2, void *stream);
/* cl[l] = 1/(2l+1) sum_m a_lm conj(b_lm) over m = -l..l for real fields (hp.alm2cl; qecl.py:148, nhl.py:175-189) */
int plk_alm2cl_dev(int lmax, const void *a, const void *b, double *cl, void *stream);
/* out_dev[0] = scale * num_dev[0] / den_dev[0]: CG step lengths (cd_solve.py:69-71, :95-99) kept on the device so that
 * the fixed-iteration multigrid stages (multigrid.py:185-215) run without host synchronisation (CUDA-graph capturable) */
int plk_scalar_ratio_dev(const double *num, const double *den, double scale, double *out, void *stream);
/* copy with change of lmax (zero fill above lmax_in) */
int plk_alm_copy_dev(int lmax_in, const void *in, int lmax_out, void *out, void *stream);
/* out (lmax_hi) = lo for l <= lsplit, hi for l > lsplit */
int plk_alm_splice_dev(int lmax_lo, const void *lo, int lmax_hi, const void *hi, int lsplit, void *out, void *stream);
/* out = lo for l <= lsplit, fl[l] * hi above: multigrid.pre_op_split with a `diag_cl` high-l branch (multigrid.py:163-182,
 * opfilt_tt.py:76-93) in one pass */
int plk_alm_splice_xfl_dev(int lmax_lo, const void *lo, int lmax_hi, const void *hi, const double *fl, int nfl, int lsplit,
                           void *out, void *stream);

/* out = ca * x + cb * y  (y may be NULL);  eblm / cd_solve vector arithmetic (util_alm.py:66-86, cd_solve.py:57-86) */
int plk_alm_lincomb_dev(long long n, double ca, const void *x, double cb, const void *y, void *out, void *stream);
/* out[l,m] = sum_{j<nterm} fl[j][l] * in[j][l,m], nterm <= 4 (host arrays of device pointers): the Wiener-filter
 * combinations of qest.py:582-588, 613-618 and the 2x2 per-l matrices of opfilt_pp.py:82-84, 101-105 */
int plk_alm_combine_dev(int lmax, int nterm, const void *const *in, const double *const *fl, const int *nfl,
                        void *out, void *stream);
/* real-harmonic packing of the dense preconditioner (qcinv/dense.py:16-53) and its mat-vec (dense.py:118-119) */
int plk_alm2rlm_dev(int lmax, const void *alm, double *rlm, void *stream);
/* same, reading the l <= lmax block of an alm stored with lmax_src >= lmax (multigrid.py:174: no intermediate alm_copy) */
int plk_alm2rlm_from_dev(int lmax, int lmax_src, const void *alm, double *rlm, void *stream);
int plk_rlm2alm_dev(int lmax, const double *rlm, void *alm, void *stream);
int plk_dense_matvec_dev(int n, const double *A, const double *x, double *y, void *stream);

/* ---- per-pixel passes (device pointers, n pixels) */
/* y = y * a            */
int plk_map_mul_dev(long long n, double *y, const double *a, void *stream);
/* result_dev[0] = sum_p a_p b_p: template_removal.py:53, :80, :107 (template_map / qmap / umap .dot) */
int plk_map_dot_dev(long long n, const double *a, const double *b, double *result_dev, void *stream);
/* QE leg products (qest.py:256-257): g *= t ; c *= t */
int plk_map_mul2_dev(long long n, double *g, double *c, const double *t, void *stream);
/* qest.py:276-278:  (re,im) = (q - i u)(g3 + i c3) - (q + i u)(g1 - i c1) */
int plk_map_qe_pp_dev(long long n, const double *q, const double *u, const double *g3, const double *c3,
                      const double *g1, const double *c1, double *re, double *im, void *stream);
/* utils_qe.py:117: d += leg_a * leg_b for complex spin maps given as (re, im) real maps; ai / bi may be NULL */
int plk_map_cmul_acc_dev(long long n, const double *ar, const double *ai, const double *br, const double *bi,
                         double *dr, double *di, void *stream);
/* opfilt_pp.py:292-301: (q,u) <- [[nqq, nqu],[nqu, nuu]] (q,u) */
int plk_map_ninv3_dev(long long n, double *q, double *u, const double *nqq, const double *nqu, const double *nuu,
                      void *stream);
/* hp.ud_grade(map, nside_out, power=-2) on RING maps: out_p = sum of the (nside_in / nside_out)^2 children of p
 * (opfilt_tt.py:172-181, opfilt_pp.py:244-251: the inverse-noise maps of the coarse multigrid levels) */
int plk_udgrade_sum_dev(int nside_in, const double *in, int nside_out, double *out, void *stream);
/* template_removal.py dot()/accum() for monopole + dipole on a RING map of the plan's nside
 * (opfilt_tt.py:193-205):
 *   if w != NULL: m_p <- m_p w_p first (in place);  sums_dev[0..3] = sum_p m_p * {1, x_p, y_p, z_p} */
int plk_map_modes_dot_dev(plk_plan *plan, double *m, const double *w, double *sums_dev, void *stream);
/*   m_p -= w_p * sum_a mode_a(p) coef_a,  coef = pinv_dev (4x4 row-major, zero rows/cols for unused modes) @ sums_dev */
int plk_map_modes_sub_dev(plk_plan *plan, double *m, const double *w, const double *sums_dev,
                          const double *pinv_dev, void *stream);

/* ---- counter-based Gaussian random numbers on the device (Philox4x32-10 + Box-Muller, FP64): the synthetic skies of
 *      the throughput runs.  Replaces the host numpy draws of plancklens/sims/phas.py:137-195 (lib_phas.get_sim,
 *      pix_lib_phas.get_sim); element i of stream `stream_id` depends on (seed, stream_id, i) only.
 *   plk_randn_dev    : out[i] = (add ? add[i] : 0) + scale * z_i, i < n   (sims/maps.py:146-173: map + nlev/vamin * phase)
 *   plk_randn_alm_dev: alm phases of a real field up to lmax: (z0 + i z1)/sqrt 2, real unit normal at m = 0
 *                      (the recipe of sims/phas.py:162-168)
 *   plk_philox_words_dev: the raw generator output, out[4 i + k], for bit-exact tests against the oracle */
int plk_randn_dev(unsigned long long seed, unsigned long long stream_id, long long n, double scale, const double *add,
                  double *out, void *stream);
int plk_randn_alm_dev(unsigned long long seed, unsigned long long stream_id, int lmax, void *alm, void *stream);
int plk_philox_words_dev(unsigned long long seed, unsigned long long stream_id, long long ncalls, unsigned int *out,
                         void *stream);

/* ---- m-partitioned ("distributed") transforms over the GPUs of one NVSwitch box: one process per GPU
 *      (BASELINE.json configs[4]: one nside-4096 transform split over 2/4/8 GPUs).
 *      The reference has no counterpart: a single healpy transform is one OpenMP process (shts.py:10).
 *
 *  Layout: the Legendre stage is split by m (blocks of `mblk` columns dealt boustrophedon to the ranks), the ring-FFT /
 *  pixel stage by ring pair (contiguous blocks of equal pixel count).  The exchange between the two stages is fused
 *  into the producing kernel: plk_dist_legendre_synth stores every phase row straight into the phase array of the
 *  rank that owns the ring pair, plk_dist_ring_anal stores column m into the array of the rank that owns m -- peer
 *  memory over NVLink (CUDA IPC mappings), no pack / all-to-all / unpack passes.  The caller provides the two
 *  barriers per transform (before the producer stage: peers are done reading; after it: all rows have landed) and,
 *  for analysis, sums the per-rank alm (every rank returns its own m rows, zero elsewhere).
 *  Maps are full-size RING arrays of which a rank reads / writes only the pixel ranges of plk_dist_pixel_ranges. */
typedef struct plk_dist plk_dist;
/* host arithmetic only (no GPU needed): pair_lo[nranks + 1] ring-pair bounds, m_owner[mmax + 1]; either may be NULL */
int plk_dist_partition(int nside, int mmax, int nranks, int mblk, int *pair_lo, int *m_owner);
int plk_dist_create(plk_dist **dist, plk_plan *plan, int rank, int nranks, int mblk /* <= 0: default 64 */);
int plk_dist_destroy(plk_dist *dist);
/* CUDA IPC handles (2 x 64 bytes) of this rank's phase arrays / mapping of a peer's */
int plk_dist_export(plk_dist *dist, void *handles128);
int plk_dist_import(plk_dist *dist, int peer, const void *handles128);
/* same-process peers (simulated ranks on one GPU in the tests) */
int plk_dist_set_peer(plk_dist *dist, int peer, void *x1, void *x2);
int plk_dist_phase_ptrs(plk_dist *dist, void **x1, void **x2);
int plk_dist_num_m(const plk_dist *dist);
/* [north lo, north hi, south lo, south hi) pixel ranges of this rank's rings */
int plk_dist_pixel_ranges(const plk_dist *dist, long long *ranges4);
int plk_dist_legendre_synth(plk_dist *dist, int spin, const void *alm1, const void *alm2, const double *fl1,
                            const double *fl2, void *stream);
int plk_dist_ring_synth(plk_dist *dist, int spin, double *map1, double *map2, void *stream);
int plk_dist_ring_anal(plk_dist *dist, int spin, const double *map1, const double *map2, void *stream);
int plk_dist_legendre_anal(plk_dist *dist, int spin, const double *fl1, const double *fl2, void *alm1, void *alm2,
                           void *stream);
/* with the additive term of plk_map2alm_add_dev on this rank's m rows (m-distributed CG forward operator: SURVEY.md
 * section 8e.2 "CG dots become a scalar all-reduce; alm stay m-distributed between calls") */
int plk_dist_legendre_anal_add(plk_dist *dist, int spin, const double *fl1, const double *fl2, const void *add1,
                               const double *afl1, const void *add2, const double *afl2, void *alm1, void *alm2,
                               void *stream);

/* ---- Wigner small-d transforms on a set of nodes x_i = cos(theta_i): what
 *      plancklens/wigners/wigners.f90:566-684 provides to utils_spin.wignerc (utils_spin.py:52-93), and through it to
 *      qresp.get_response and nhl.get_nhl.  All pointers are device pointers.
 *   wignerpos  : xi[i] = sum_l cl[l] (2l+1)/(4 pi) d^l_{s1 s2}(x[i])
 *   wignercoeff: cl[l] = 2 pi sum_i f[i] d^l_{s1 s2}(x[i]),  0 <= l <= lmax   (f = function values times quadrature weights) */
int plk_wignerpos_dev(const double *cl, int lmax, const double *x, int nx, int s1, int s2, double *xi, void *stream);
int plk_wignercoeff_dev(const double *f, const double *x, int nx, int s1, int s2, int lmax, double *cl, void *stream);

/* ---- measurement helpers (bench.py): per-kernel CUDA-event timing of the Legendre launches and the device's
 *      FP64 FMA peak.  kinds: 0 synthesis spin 0, 1 synthesis spin s, 2 analysis spin 0, 3 analysis spin s,
 *      4 synthesis spin s with zero curl input (gradient-only kernel: 16 instead of 24 flop per unit). */
int plk_profile_enable(int on);
int plk_profile_read(int *counts5, double *total_ms5);
int plk_fp64_peak(double *tflops, int reps);
/* share of the (l, m, ring pair) volume the Legendre kernels walk for this spin (the rest lies below the 2^-60
 * start threshold near the poles and is skipped) */
int plk_plan_active_fraction(plk_plan *plan, int spin, double *frac);

#ifdef __cplusplus
}
#endif
#endif /* PLK_H */
