// libplk_b200: plan management and C ABI (see include/plk.h).  sm_100a only.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <map>
#include <numeric>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/plk.h"
#include "plk_blas.cuh"
#include "plk_common.h"
#include "plk_fft.cuh"
#include "plk_legendre.cuh"
#include "plk_rng.cuh"
#include "plk_tables.h"
#include "plk_wigner.cuh"

using namespace plk;

// ------------------------------------------------------------------------------------------ errors / counters
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) return fail(PLK_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
#define LAUNCHED()                                                                                 \
  do {                                                                                             \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                            \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess) return fail(PLK_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// optional per-kernel timing of the Legendre launches (CUDA events on the launching stream)
struct ProfRec { cudaEvent_t e0, e1; int kind; };   // kind: 0 synth spin0, 1 synth spin s, 2 anal spin0, 3 anal spin s,
                                                    //       4 synth spin s gradient-only
static bool g_prof = false;
static std::vector<ProfRec> g_prof_recs;
static void prof_begin(int kind, cudaStream_t st) {
  if (!g_prof) return;
  ProfRec r; r.kind = kind;
  cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
  g_prof_recs.push_back(r);
}
static void prof_end(cudaStream_t st) {
  if (!g_prof) return;
  cudaEventRecord(g_prof_recs.back().e1, st);
}

extern "C" const char *plk_last_error(void) { return g_err.c_str(); }
extern "C" int plk_version(void) { return 100; }
extern "C" long long plk_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------ plan
struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
};

struct SpinDev {
  bool ready = false;
  DevSpin d{};
  std::vector<void *> owned;
};

struct plk_plan {
  int nside, lmax, mmax, npair, nring, device;
  long long npix;
  int pitch;  // phase array row pitch (complex elements)
  HostGeom hg;
  DevGeom g{};
  DevFFT f{};
  DevRings rings{};
  int fft_smem = 0;
  struct FftClass { int M, nbatch, offset, count, smem, threads; };
  std::vector<FftClass> fft_classes;
  int *fft_list = nullptr;      // ring pairs grouped by FFT size class
  int *fft_order = nullptr;
  int *morder = nullptr;
  std::vector<void *> owned;
  SpinDev spins[4];
  int seed_thr_exp = kSeedThrExp;
  DevBuf X1, X2, rec, part, partial, hostio[4];
  long long table_bytes = 0;
};

// m-partitioned execution of one plan over the GPUs of a box (include/plk.h, "distributed transforms")
struct plk_dist {
  plk_plan *plan = nullptr;
  int rank = 0, nranks = 1, mblk = 64;
  std::vector<int> pair_lo;          // [nranks + 1]
  std::vector<int> mlist;            // local m, longest recurrences first
  int *d_mlist = nullptr;
  int *d_fft_list = nullptr;         // local ring pairs grouped by FFT size class
  std::vector<plk_plan::FftClass> classes;
  cplx *X1 = nullptr, *X2 = nullptr; // own exchange buffers: plain cudaMalloc so that CUDA IPC can export them.
                                     // analysis: phase arrays [ring][pitch]; synthesis: transposed, [m][tpitch]
  int tpitch = 0;
  cplx *px1[kMaxRanks] = {}, *px2[kMaxRanks] = {};
  bool opened[kMaxRanks] = {};
};

template <class T>
static int upload(plk_plan *p, const std::vector<T> &h, T **d, std::vector<void *> *owned = nullptr) {
  const size_t b = std::max<size_t>(h.size() * sizeof(T), 16);
  CK(cudaMalloc((void **)d, b));
  (owned ? *owned : p->owned).push_back(*d);
  p->table_bytes += b;
  if (!h.empty()) CK(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}
static int ensure(DevBuf &b, size_t bytes) {
  if (b.bytes >= bytes) return 0;
  if (b.p) cudaFree(b.p);
  b.p = nullptr; b.bytes = 0;
  cudaError_t e = cudaMalloc(&b.p, bytes);
  if (e != cudaSuccess) return fail(PLK_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  b.bytes = bytes;
  return 0;
}

static int nextpow2(int v) { int r = 1; while (r < v) r <<= 1; return r; }
static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v ? atoi(v) : dflt;
}

extern "C" int plk_plan_create(plk_plan **out, int nside, int lmax, int mmax) {
  if (!out) return fail(PLK_EINVAL, "plan pointer is NULL");
  *out = nullptr;
  if (nside < 1 || nside > 8192 || (nside & (nside - 1))) return fail(PLK_EINVAL, "nside must be a power of two in [1, 8192], got %d", nside);
  if (lmax < 0 || lmax > 16384) return fail(PLK_EINVAL, "lmax out of range: %d", lmax);
  if (mmax != lmax) return fail(PLK_EINVAL, "mmax (%d) must equal lmax (%d)", mmax, lmax);
  int dev;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(PLK_ENODEV, "no CUDA device: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(PLK_ENODEV, "libplk_b200 is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);

  plk_plan *p = new plk_plan();
  p->nside = nside; p->lmax = lmax; p->mmax = mmax; p->device = dev;
  p->hg = make_geom(nside);
  const HostGeom &hg = p->hg;
  p->npair = hg.npair; p->nring = hg.nring; p->npix = hg.npix;
  p->pitch = (mmax + 1 + 1) & ~1;   // even: 32-byte aligned rows

  // geometry
  DevGeom &g = p->g;
  g.nside = nside; g.npair = hg.npair; g.nring = hg.nring; g.npix = hg.npix;
  double *d;
#define UP(vec, field) do { int rc = upload(p, vec, &d); if (rc) { plk_plan_destroy(p); return rc; } field = d; } while (0)
  UP(hg.cth, g.cth); UP(hg.sh_hi, g.sh_hi); UP(hg.sh_lo, g.sh_lo); UP(hg.ch_hi, g.ch_hi); UP(hg.ch_lo, g.ch_lo);
#undef UP

  // ring FFT tables
  DevFFT &f = p->f;
  f.nside = nside; f.npair = hg.npair; f.nring = hg.nring;
  std::vector<int> M(hg.npair), shifted = hg.shifted, nphi = hg.nphi;
  std::vector<long long> voff(hg.npair, -1), sn(hg.npair), ss(hg.npair);
  long long vtot = 0;
  int Mmax = 16;
  std::vector<double> cost(hg.npair);
  for (int ip = 0; ip < hg.npair; ++ip) {
    const int q = hg.nphi[ip] / 4;
    sn[ip] = hg.start_n[ip]; ss[ip] = hg.start_s[ip];
    if (q <= kTinyQ) { M[ip] = 0; cost[ip] = 4.0 * q * (mmax + 1) * 20; }
    else if ((q & (q - 1)) == 0) { M[ip] = q; cost[ip] = 4.0 * q * ilog2(q); }
    else { M[ip] = nextpow2(2 * q - 1); voff[ip] = vtot; vtot += M[ip]; cost[ip] = 4.0 * 2.2 * M[ip] * ilog2(M[ip]); }
    Mmax = std::max(Mmax, M[ip]);
  }
  f.Wn = Mmax;
  std::vector<cplx> W(Mmax);
  for (int k = 0; k < Mmax; ++k) {
    long double a = -2.0L * 3.14159265358979323846264338327950288L * k / Mmax;
    W[k] = mk((double)cosl(a), (double)sinl(a));
  }
  {
    cplx *dW; int rc = upload(p, W, &dW); if (rc) { plk_plan_destroy(p); return rc; } f.W = dW;
    int *di; long long *dl;
    rc = upload(p, M, &di); if (rc) { plk_plan_destroy(p); return rc; } f.M = di;
    rc = upload(p, nphi, &di); if (rc) { plk_plan_destroy(p); return rc; } f.nphi = di;
    rc = upload(p, shifted, &di); if (rc) { plk_plan_destroy(p); return rc; } f.shifted = di;
    rc = upload(p, voff, &dl); if (rc) { plk_plan_destroy(p); return rc; } f.voff = dl;
    rc = upload(p, sn, &dl); if (rc) { plk_plan_destroy(p); return rc; } f.start_n = dl;
    rc = upload(p, ss, &dl); if (rc) { plk_plan_destroy(p); return rc; } f.start_s = dl;
    std::vector<int> order(hg.npair);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    rc = upload(p, order, &di); if (rc) { plk_plan_destroy(p); return rc; } p->fft_order = di; f.order = di;
    std::vector<int> mo(mmax + 1);
    std::iota(mo.begin(), mo.end(), 0);
    rc = upload(p, mo, &di); if (rc) { plk_plan_destroy(p); return rc; } p->morder = di;
  }
  // size classes: one launch per distinct FFT size so that every block of a launch needs the same shared memory
  {
    std::map<int, std::vector<int>> by_m;
    for (int ip = 0; ip < hg.npair; ++ip) by_m[M[ip]].push_back(ip);
    std::vector<int> list;
    const int nb4_maxm = env_int("PLK_FFT_NB4_MAXM", 1024);
    const int nb1_minm = env_int("PLK_FFT_NB1_MINM", 8192);   // M = 8192 (nside 4096 caps): one DFT buffer only
    f.nb4_maxm = nb4_maxm; f.nb1_minm = nb1_minm;
    for (auto it = by_m.rbegin(); it != by_m.rend(); ++it) {     // largest transforms first
      plk_plan::FftClass c;
      c.M = it->first; c.offset = (int)list.size(); c.count = (int)it->second.size();
      c.nbatch = c.M == 0 ? 0 : (c.M >= nb1_minm ? 1 : (c.M <= nb4_maxm ? 4 : 2));
      c.smem = c.M == 0 ? 0 : (c.nbatch * c.M + c.M / 4) * (int)sizeof(cplx);
      c.threads = (c.nbatch * c.M / 16 >= 512) ? 512 : 256;   // one radix-16 butterfly per thread and pass
      p->fft_smem = std::max(p->fft_smem, c.smem);
      list.insert(list.end(), it->second.begin(), it->second.end());
      p->fft_classes.push_back(c);
    }
    int *di; int rc = upload(p, list, &di); if (rc) { plk_plan_destroy(p); return rc; } p->fft_list = di;
    // Small plans (the coarse multigrid levels of the qcinv chains): one launch over all classes with the largest
    // shared-memory request -- each per-class launch holds a handful of blocks for ~20 us, so six of them back to back
    // cost more than the transform itself.
    if (hg.npair <= env_int("PLK_FFT_MERGE_MAXPAIR", 1024) && p->fft_classes.size() > 1) {
      plk_plan::FftClass c;
      c.M = -1; c.nbatch = 0; c.offset = 0; c.count = hg.npair; c.smem = p->fft_smem; c.threads = 256;
      p->fft_classes.assign(1, c);
    } else if (env_int("PLK_FFT_MERGE_256", 1)) {
      // Large plans: all 256-thread classes (M <= 2048) ask for at most ~74 KB, i.e. the same three blocks per SM as
      // the largest of them, so they run as ONE launch.  The near-polar classes hold a handful of ring pairs each and
      // have long serial tails (aliasing folds of mmax / n terms per bin); launched one after the other they cost
      // 50 - 110 us apiece, inside the big launch they hide behind the equatorial rings.
      std::vector<plk_plan::FftClass> out;
      plk_plan::FftClass m;
      m.M = -1; m.nbatch = 0; m.offset = -1; m.count = 0; m.smem = 0; m.threads = 256;
      for (const auto &c : p->fft_classes) {
        if (c.threads != 256) { out.push_back(c); continue; }
        if (m.offset < 0) m.offset = c.offset;
        m.count += c.count;
        m.smem = std::max(m.smem, c.smem);
      }
      if (m.count > 0) out.push_back(m);
      p->fft_classes = out;
    }
  }
  if (p->fft_smem > 227 * 1024) { plk_plan_destroy(p); return fail(PLK_EINVAL, "ring FFT needs %d bytes of shared memory", p->fft_smem); }
  {
    static int attr_smem = 0;   // the attribute is per function, not per plan: only ever raise it
    if (p->fft_smem > attr_smem) {
      cudaError_t e1 = cudaFuncSetAttribute(ring_synth_kernel<256, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->fft_smem);
      cudaError_t e2 = cudaFuncSetAttribute(ring_anal_kernel<256, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->fft_smem);
      cudaError_t e3 = cudaFuncSetAttribute(bluestein_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p->fft_smem);
      cudaError_t e4 = cudaFuncSetAttribute(ring_synth_kernel<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->fft_smem);
      cudaError_t e5 = cudaFuncSetAttribute(ring_anal_kernel<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->fft_smem);
      if (e4 != cudaSuccess || e5 != cudaSuccess) e1 = e4 != cudaSuccess ? e4 : e5;
      if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { plk_plan_destroy(p); return fail(PLK_ECUDA, "cudaFuncSetAttribute(smem=%d) failed", p->fft_smem); }
      attr_smem = p->fft_smem;
    }
  }
  if (vtot > 0) {
    cplx *dV;
    cudaError_t ee = cudaMalloc((void **)&dV, (size_t)vtot * sizeof(cplx));
    if (ee != cudaSuccess) { plk_plan_destroy(p); return fail(PLK_ENOMEM, "cudaMalloc(V) failed"); }
    p->owned.push_back(dV); p->table_bytes += vtot * sizeof(cplx);
    f.V = dV;
    bluestein_setup_kernel<<<hg.npair, kFftThreads, p->fft_smem>>>(f, dV);
    g_launches.fetch_add(1);
    ee = cudaDeviceSynchronize();
    if (ee != cudaSuccess) { plk_plan_destroy(p); return fail(PLK_ECUDA, "bluestein setup failed: %s", cudaGetErrorString(ee)); }
  } else {
    f.V = nullptr;
  }
  f.mtop = nullptr;
  f.dist_n = 0; f.dist_mblk = 1;
  f.pix.n = 0;

  // per-ring table for the template (monopole / dipole) kernels
  {
    std::vector<long long> st(hg.nring);
    std::vector<int> np(hg.nring), sh(hg.nring);
    std::vector<double> z(hg.nring), sth(hg.nring);
    for (int ip = 0; ip < hg.npair; ++ip) {
      const int rn = ip, rs = hg.nring - 1 - ip;
      st[rn] = hg.start_n[ip]; np[rn] = hg.nphi[ip]; sh[rn] = hg.shifted[ip]; z[rn] = hg.cth[ip]; sth[rn] = hg.sth[ip];
      if (rs != rn) { st[rs] = hg.start_s[ip]; np[rs] = hg.nphi[ip]; sh[rs] = hg.shifted[ip]; z[rs] = -hg.cth[ip]; sth[rs] = hg.sth[ip]; }
    }
    long long *dl; int *di; double *dd;
    int rc;
    rc = upload(p, st, &dl); if (rc) { plk_plan_destroy(p); return rc; } p->rings.start = dl;
    rc = upload(p, np, &di); if (rc) { plk_plan_destroy(p); return rc; } p->rings.nphi = di;
    rc = upload(p, sh, &di); if (rc) { plk_plan_destroy(p); return rc; } p->rings.shifted = di;
    rc = upload(p, z, &dd); if (rc) { plk_plan_destroy(p); return rc; } p->rings.z = dd;
    rc = upload(p, sth, &dd); if (rc) { plk_plan_destroy(p); return rc; } p->rings.sth = dd;
    p->rings.nring = hg.nring;
  }
  *out = p;
  return PLK_OK;
}

extern "C" int plk_plan_destroy(plk_plan *p) {
  if (!p) return PLK_OK;
  for (void *q : p->owned) cudaFree(q);
  for (auto &s : p->spins) for (void *q : s.owned) cudaFree(q);
  for (DevBuf *b : {&p->X1, &p->X2, &p->rec, &p->part, &p->partial}) if (b->p) cudaFree(b->p);
  for (auto &b : p->hostio) if (b.p) cudaFree(b.p);
  delete p;
  return PLK_OK;
}

extern "C" long long plk_plan_device_bytes(const plk_plan *p) {
  if (!p) return 0;
  long long b = p->table_bytes;
  for (const DevBuf *x : {&p->X1, &p->X2, &p->rec, &p->part, &p->partial}) b += (long long)x->bytes;
  for (auto &x : p->hostio) b += (long long)x.bytes;
  return b;
}
// Start threshold of the Legendre recurrences (2^exp2): drops the per-spin tables so that they are rebuilt with the
// new value on next use.  Not to be called while a captured CUDA graph still refers to this plan's tables.
extern "C" int plk_plan_set_seed_threshold(plk_plan *p, int exp2) {
  if (!p) return fail(PLK_EINVAL, "plan is NULL");
  if (exp2 > -20 || exp2 < -900) return fail(PLK_EINVAL, "threshold exponent %d outside [-900, -20]", exp2);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(PLK_ECUDA, "cudaDeviceSynchronize failed: %s", cudaGetErrorString(e));
  for (auto &s : p->spins) {
    for (void *q : s.owned) cudaFree(q);
    s.owned.clear();
    s.ready = false;
  }
  p->seed_thr_exp = exp2;
  return PLK_OK;
}
extern "C" int plk_plan_nside(const plk_plan *p) { return p ? p->nside : 0; }
extern "C" int plk_plan_lmax(const plk_plan *p) { return p ? p->lmax : 0; }

static int leg_attrs();
// per-spin recurrence tables + seeds (built on first use)
static int ensure_spin(plk_plan *p, int spin) {
  if (spin < 0 || spin > 3) return fail(PLK_EINVAL, "spin must be 0..3, got %d", spin);
  SpinDev &sd = p->spins[spin];
  if (sd.ready) return 0;
  { int rca = leg_attrs(); if (rca) return rca; }
  SpinTables t = make_spin_tables(spin, p->lmax, p->mmax);
  const size_t n = t.U.size();
  std::vector<double2> uv(n);
  for (size_t i = 0; i < n; ++i) uv[i] = make_double2(t.U[i], t.V[i]);
  DevSpin &d = sd.d;
  d.spin = spin; d.lmax = p->lmax; d.mmax = p->mmax; d.thr_exp = p->seed_thr_exp;
  int rc;
  double2 *duv; double *dd; int *di; signed char *dc;
#define UPS(vec, ptr, field) do { rc = upload(p, vec, &ptr, &sd.owned); if (rc) return rc; field = ptr; } while (0)
  UPS(uv, duv, d.UV); UPS(t.alpha, dd, d.alpha); UPS(t.k_hi, dd, d.k_hi); UPS(t.k_lo, dd, d.k_lo);
  UPS(t.k_e, di, d.k_e); UPS(t.pc, di, d.pc); UPS(t.ps, di, d.ps); UPS(t.sg_p, dc, d.sg_p); UPS(t.sg_m, dc, d.sg_m);
#undef UPS
  const size_t ns = (size_t)(p->mmax + 1) * p->npair;
  auto dalloc = [&](void **q, size_t bytes) -> int {
    cudaError_t e = cudaMalloc(q, bytes);
    if (e != cudaSuccess) return fail(PLK_ENOMEM, "cudaMalloc(seeds, %zu) failed", bytes);
    sd.owned.push_back(*q); p->table_bytes += bytes;
    return 0;
  };
  if ((rc = dalloc((void **)&d.ks, ns * sizeof(int)))) return rc;
  if ((rc = dalloc((void **)&d.s0, ns * sizeof(double)))) return rc;
  if ((rc = dalloc((void **)&d.s1, ns * sizeof(double)))) return rc;
  d.s2 = d.s3 = nullptr;
  if (spin > 0) {
    if ((rc = dalloc((void **)&d.s2, ns * sizeof(double)))) return rc;
    if ((rc = dalloc((void **)&d.s3, ns * sizeof(double)))) return rc;
  }
  dim3 grid((p->npair + 127) / 128, p->mmax + 1);
  if (spin == 0) seed_kernel<false><<<grid, 128>>>(p->g, d);
  else seed_kernel<true><<<grid, 128>>>(p->g, d);
  LAUNCHED();
  if ((rc = dalloc((void **)&d.mtop, (size_t)p->npair * sizeof(int)))) return rc;
  mtop_kernel<<<(p->npair + 127) / 128, 128>>>(p->g, d);
  LAUNCHED();
  CK(cudaDeviceSynchronize());
  sd.ready = true;
  return 0;
}

// ------------------------------------------------------------------------------------------ Legendre launches
template <bool SPIN, bool SYNTH>
static size_t leg_smem(int nr) {
  size_t s = 128 + (size_t)kStages * StageBytes<SPIN, SYNTH>::stage;
  if (!SYNTH) s += (size_t)2 * kNCW * kChunk * (SPIN ? 4 : 2) * sizeof(double);
  if (SYNTH || !SPIN) s += (size_t)nr * (SPIN ? 4 : 2) * kNCW * 32 * sizeof(double);     // seeds
  return s;
}
// Every variant fits the default 48 KB dynamic shared memory limit.  Do NOT raise
// cudaFuncAttributeMaxDynamicSharedMemorySize "just in case": it moves the L1 / shared carve-out and slowed the
// spin-s analysis kernel (divergent seed loads living in L1) from 10.4 to 11.9 ms on B200.
static int leg_attrs() { return 0; }
static int pick_nr(const plk_plan *p, int nrmax) {
  // enough blocks to cover the SMs a few times over; small grids get fewer pairs per thread
  int nr = nrmax;
  while (nr > 1) {
    const long long tiles = (p->npair + kNCW * 32 * nr - 1) / (kNCW * 32 * nr);
    if (tiles * (p->mmax + 1) >= 4 * 148 && p->npair >= kNCW * 32 * nr) break;
    nr >>= 1;
  }
  return nr;
}

// Work buffers have ONE size per plan (the largest any spin needs): pointers baked into captured CUDA graphs
// (qcinv multigrid stages) must never be invalidated by a later call with another spin.
static int ensure_work(plk_plan *p) {
  const size_t nalm = (size_t)alm_size(p->lmax, p->mmax);
  int rc;
  if ((rc = ensure(p->rec, nalm * 32 + 64))) return rc;
  size_t pb = 0;
  for (int sp = 0; sp < 2; ++sp) {
    const int nr = pick_nr(p, sp ? env_int("PLK_NR_ANAS", 4) : env_int("PLK_NR_ANA0", 4));
    const size_t ntile = (p->npair + kNCW * 32 * nr - 1) / (kNCW * 32 * nr);
    pb = std::max(pb, ntile * nalm * (sp ? 4 : 2) * sizeof(double));
  }
  if ((rc = ensure(p->part, pb))) return rc;
  const size_t b = (size_t)p->nring * p->pitch * sizeof(cplx);
  if ((rc = ensure(p->X1, b)) || (rc = ensure(p->X2, b))) return rc;
  return 0;
}

static DistX no_dist() { DistX d; memset(&d, 0, sizeof d); d.nranks = 1; d.mblk = 1; return d; }
static DistX dist_x(const plk_dist *d) {
  DistX x = no_dist();
  x.nranks = d->nranks; x.mblk = d->mblk; x.tpitch = d->tpitch;
  for (int q = 0; q <= d->nranks; ++q) x.pair_lo[q] = d->pair_lo[q];
  for (int q = 0; q < d->nranks; ++q) { x.x1[q] = d->px1[q]; x.x2[q] = d->px2[q]; }
  return x;
}

static int legendre_synth(plk_plan *p, int spin, const void *alm1, const void *alm2, const double *fl1,
                          const double *fl2, cplx *X1, cplx *X2, cudaStream_t st, const plk_dist *dd = nullptr) {
  int rc = ensure_spin(p, spin);
  if (rc) return rc;
  const DevSpin &d = p->spins[spin].d;
  if ((rc = ensure_work(p))) return rc;
  const int nm = dd ? (int)dd->mlist.size() : p->mmax + 1;        // m columns handled by this process
  const int *morder = dd ? dd->d_mlist : p->morder;
  const DistX dx = dd ? dist_x(dd) : no_dist();
  if (nm == 0) return 0;
  dim3 pg((p->lmax + 256) / 256, nm);
  if (spin == 0) prep_alm_kernel<false><<<pg, 256, 0, st>>>(d, (const cplx *)alm1, nullptr, fl1, nullptr, p->rec.p, morder);
  else prep_alm_kernel<true><<<pg, 256, 0, st>>>(d, (const cplx *)alm1, (const cplx *)alm2, fl1, fl2, p->rec.p, morder);
  LAUNCHED();
  const int nr = pick_nr(p, spin ? env_int("PLK_NR_SYNS", 4) : env_int("PLK_NR_SYN0", 4));
  dim3 grid((p->npair + kNCW * 32 * nr - 1) / (kNCW * 32 * nr), nm);
  const int nthr = kLegThreads;
#define SYN(SP, NR)                                                                                              \
  legendre_synth_kernel<SP, NR><<<grid, nthr, leg_smem<SP, true>(NR), st>>>(p->g, d, p->rec.p, X1, X2, p->pitch, morder, dx)
#define SYNG(NR)                                                                                                  \
  legendre_synth_kernel<true, NR, true><<<grid, nthr, leg_smem<true, true>(NR), st>>>(p->g, d, p->rec.p, X1, X2, p->pitch, morder, dx)
  const bool grad = spin > 0 && !alm2 && env_int("PLK_SYN_GRAD", 1);
  prof_begin(spin ? (grad ? 4 : 1) : 0, st);
  if (spin == 0) {
    if (nr == 4) SYN(false, 4); else if (nr == 2) SYN(false, 2); else SYN(false, 1);
  } else if (grad) {                                      // zero curl component: gradient-only kernel
    if (nr == 4) SYNG(4); else if (nr == 2) SYNG(2); else SYNG(1);
  } else {
    if (nr == 4) SYN(true, 4); else if (nr == 2) SYN(true, 2); else SYN(true, 1);
  }
#undef SYN
#undef SYNG
  prof_end(st);
  LAUNCHED();
  return 0;
}

struct AlmAdd { const cplx *a1 = nullptr, *a2 = nullptr; const double *f1 = nullptr, *f2 = nullptr; };
static int legendre_anal(plk_plan *p, int spin, const cplx *X1, const cplx *X2, const double *fl1, const double *fl2,
                         void *alm1, void *alm2, cudaStream_t st, const plk_dist *dd = nullptr, AlmAdd add = AlmAdd()) {
  int rc = ensure_spin(p, spin);
  if (rc) return rc;
  const DevSpin &d = p->spins[spin].d;
  const size_t nalm = (size_t)alm_size(p->lmax, p->mmax);
  const int nv = spin ? 4 : 2;
  const int nr = pick_nr(p, spin ? env_int("PLK_NR_ANAS", 4) : env_int("PLK_NR_ANA0", 4));
  const int ntile = (p->npair + kNCW * 32 * nr - 1) / (kNCW * 32 * nr);
  const long long stride = (long long)nalm * nv;
  if ((rc = ensure_work(p))) return rc;
  const int nm = dd ? (int)dd->mlist.size() : p->mmax + 1;
  const int *morder = dd ? dd->d_mlist : p->morder;
  if (dd) {   // rows of other ranks stay zero: the caller sums the per-rank alms (or keeps them m-distributed)
    CK(cudaMemsetAsync(alm1, 0, nalm * sizeof(cplx), st));
    if (spin) CK(cudaMemsetAsync(alm2, 0, nalm * sizeof(cplx), st));
  }
  if (nm == 0) return 0;
  dim3 grid(ntile, nm);
  const int nthr = kLegThreads;
#define ANA(SP, NR)                                                                                              \
  legendre_anal_kernel<SP, NR><<<grid, nthr, leg_smem<SP, false>(NR), st>>>(p->g, d, X1, X2, p->pitch, (double *)p->part.p, stride, morder, env_int("PLK_DBG_ANA", 0))
  prof_begin(spin ? 3 : 2, st);
  if (spin == 0) {
    if (nr == 4) ANA(false, 4); else if (nr == 2) ANA(false, 2); else ANA(false, 1);
  } else {
    if (nr == 4) ANA(true, 4); else if (nr == 2) ANA(true, 2); else ANA(true, 1);
  }
#undef ANA
  prof_end(st);
  LAUNCHED();
  dim3 pg((p->lmax + 256) / 256, nm);
  if (spin == 0) finish_alm_kernel<false><<<pg, 256, 0, st>>>(d, (const double *)p->part.p, stride, ntile, fl1, nullptr, (cplx *)alm1, nullptr, morder, add.a1, add.f1, nullptr, nullptr);
  else finish_alm_kernel<true><<<pg, 256, 0, st>>>(d, (const double *)p->part.p, stride, ntile, fl1, fl2, (cplx *)alm1, (cplx *)alm2, morder, add.a1, add.f1, add.a2, add.f2);
  LAUNCHED();
  return 0;
}

// dd != null: only this rank's ring pairs; analysis scatters column m to the rank owning it (which = 0: X1, 1: X2)
static int ring_synth(plk_plan *p, const cplx *X, double *map, cudaStream_t st, const int *mtop = nullptr,
                      const plk_dist *dd = nullptr) {
  DevFFT f = p->f;
  f.mtop = mtop;
  const int *list = dd ? dd->d_fft_list : p->fft_list;
  for (const auto &c : (dd ? dd->classes : p->fft_classes)) {
    if (c.count == 0) continue;
    if (c.threads == 512) ring_synth_kernel<512, 1><<<c.count, 512, c.smem, st>>>(f, list + c.offset, c.nbatch, X, p->pitch, p->mmax, map);
    else ring_synth_kernel<256, 3><<<c.count, 256, c.smem, st>>>(f, list + c.offset, c.nbatch, X, p->pitch, p->mmax, map);
    LAUNCHED();
  }
  return 0;
}
static int ring_anal(plk_plan *p, const double *map, cplx *X, cudaStream_t st, const int *mtop = nullptr,
                     const plk_dist *dd = nullptr, int which = 0, const PixProg *pix = nullptr) {
  const double w = 4.0 * M_PI / (double)p->npix;
  DevFFT f = p->f;
  f.mtop = mtop;
  if (pix) f.pix = *pix;
  if (dd) {
    f.dist_n = dd->nranks; f.dist_mblk = dd->mblk;
    for (int q = 0; q < dd->nranks; ++q) f.dist_x[q] = which ? dd->px2[q] : dd->px1[q];
  }
  const int *list = dd ? dd->d_fft_list : p->fft_list;
  for (const auto &c : (dd ? dd->classes : p->fft_classes)) {
    if (c.count == 0) continue;
    if (c.threads == 512) ring_anal_kernel<512, 1><<<c.count, 512, c.smem, st>>>(f, list + c.offset, c.nbatch, map, X, p->pitch, p->mmax, w);
    else ring_anal_kernel<256, 3><<<c.count, 256, c.smem, st>>>(f, list + c.offset, c.nbatch, map, X, p->pitch, p->mmax, w);
    LAUNCHED();
  }
  return 0;
}
static int ensure_phase(plk_plan *p, int /*ncomp*/) { return ensure_work(p); }

#define CHECK_PLAN(p) do { if (!(p)) return fail(PLK_EINVAL, "plan is NULL"); } while (0)

extern "C" int plk_legendre_synth_dev(plk_plan *p, int spin, const void *alm1, const void *alm2, const double *fl1,
                                      const double *fl2, void *X1, void *X2, void *stream) {
  CHECK_PLAN(p);
  if (!alm1 || !X1 || (spin > 0 && !X2)) return fail(PLK_EINVAL, "NULL buffer");
  return legendre_synth(p, spin, alm1, alm2, fl1, fl2, (cplx *)X1, (cplx *)X2, (cudaStream_t)stream);
}
extern "C" int plk_legendre_anal_dev(plk_plan *p, int spin, const void *X1, const void *X2, const double *fl1,
                                     const double *fl2, void *alm1, void *alm2, void *stream) {
  CHECK_PLAN(p);
  if (!alm1 || !X1 || (spin > 0 && (!X2 || !alm2))) return fail(PLK_EINVAL, "NULL buffer");
  return legendre_anal(p, spin, (const cplx *)X1, (const cplx *)X2, fl1, fl2, alm1, alm2, (cudaStream_t)stream);
}
extern "C" int plk_ring_synth_dev(plk_plan *p, const void *X, double *map, void *stream) {
  CHECK_PLAN(p);
  if (!X || !map) return fail(PLK_EINVAL, "NULL buffer");
  return ring_synth(p, (const cplx *)X, map, (cudaStream_t)stream);
}
extern "C" int plk_ring_anal_dev(plk_plan *p, const double *map, void *X, void *stream) {
  CHECK_PLAN(p);
  if (!X || !map) return fail(PLK_EINVAL, "NULL buffer");
  return ring_anal(p, map, (cplx *)X, (cudaStream_t)stream);
}

extern "C" int plk_alm2map_dev(plk_plan *p, int spin, const void *alm1, const void *alm2, const double *fl1,
                               const double *fl2, double *map1, double *map2, void *stream) {
  CHECK_PLAN(p);
  if (spin < 0 || spin > 3) return fail(PLK_EINVAL, "spin must be 0..3, got %d", spin);
  if (!alm1 || !map1 || (spin > 0 && !map2)) return fail(PLK_EINVAL, "NULL buffer");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_phase(p, spin ? 2 : 1);
  if (rc) return rc;
  if ((rc = legendre_synth(p, spin, alm1, alm2, fl1, fl2, (cplx *)p->X1.p, (cplx *)p->X2.p, st))) return rc;
  const int *mtop = p->spins[spin].d.mtop;
  if ((rc = ring_synth(p, (const cplx *)p->X1.p, map1, st, mtop))) return rc;
  if (spin > 0 && (rc = ring_synth(p, (const cplx *)p->X2.p, map2, st, mtop))) return rc;
  return PLK_OK;
}

static int to_pixprog(const plk_pixprog *in, long long npix, PixProg *out) {
  if (in->nterm < 1 || in->nterm > kMaxPixTerms) return fail(PLK_EINVAL, "pixel program needs 1..%d terms, got %d", kMaxPixTerms, in->nterm);
  out->n = in->nterm;
  for (int k = 0; k < kMaxPixTerms; ++k) { out->a[k] = nullptr; out->b[k] = nullptr; out->s[k] = 0.0; }
  for (int k = 0; k < in->nterm; ++k) {
    if (!in->a[k]) return fail(PLK_EINVAL, "pixel program term %d has no map", k);
    if (((uintptr_t)in->a[k] & 31) || ((uintptr_t)in->b[k] & 31)) return fail(PLK_EINVAL, "pixel program maps must be 32-byte aligned");
    out->a[k] = in->a[k]; out->b[k] = in->b[k]; out->s[k] = in->scale[k];
  }
  (void)npix;
  return 0;
}
static int map2alm_impl(plk_plan *p, int spin, const double *map1, const double *map2, const double *fl1,
                        const double *fl2, void *alm1, void *alm2, void *stream, AlmAdd add,
                        const plk_pixprog *pix1 = nullptr, const plk_pixprog *pix2 = nullptr) {
  CHECK_PLAN(p);
  if (spin < 0 || spin > 3) return fail(PLK_EINVAL, "spin must be 0..3, got %d", spin);
  if (!alm1 || (!map1 && !pix1) || (spin > 0 && ((!map2 && !pix2) || !alm2))) return fail(PLK_EINVAL, "NULL buffer");
  PixProg q1, q2;
  q1.n = q2.n = 0;
  if (pix1) { int rc0 = to_pixprog(pix1, p->npix, &q1); if (rc0) return rc0; }
  if (spin > 0 && pix2) { int rc0 = to_pixprog(pix2, p->npix, &q2); if (rc0) return rc0; }
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_phase(p, spin ? 2 : 1);
  if (rc) return rc;
  if ((rc = ensure_spin(p, spin))) return rc;
  const int *mtop = p->spins[spin].d.mtop;
  if ((rc = ring_anal(p, map1, (cplx *)p->X1.p, st, mtop, nullptr, 0, pix1 ? &q1 : nullptr))) return rc;
  if (spin > 0 && (rc = ring_anal(p, map2, (cplx *)p->X2.p, st, mtop, nullptr, 0, pix2 ? &q2 : nullptr))) return rc;
  return legendre_anal(p, spin, (const cplx *)p->X1.p, (const cplx *)p->X2.p, fl1, fl2, alm1, alm2, st, nullptr, add);
}
// Analysis of maps that are never materialised: component c of the input is the pixel program pix_c (see plk_pixprog);
// optional additive term as in plk_map2alm_add_dev (add1 == NULL: none).
extern "C" int plk_map2alm_pix_dev(plk_plan *p, int spin, const plk_pixprog *pix1, const plk_pixprog *pix2, const double *fl1,
                                   const double *fl2, const void *add1, const double *afl1, const void *add2,
                                   const double *afl2, void *alm1, void *alm2, void *stream) {
  if (!pix1 || (spin > 0 && !pix2)) return fail(PLK_EINVAL, "NULL pixel program");
  AlmAdd add;
  if (add1) {
    if (!afl1 || (spin > 0 && (!add2 || !afl2))) return fail(PLK_EINVAL, "NULL additive term");
    if (add1 == alm1 || (spin > 0 && add2 == alm2)) return fail(PLK_EINVAL, "additive term aliases the output");
    add.a1 = (const cplx *)add1; add.f1 = afl1; add.a2 = (const cplx *)add2; add.f2 = afl2;
  }
  return map2alm_impl(p, spin, nullptr, nullptr, fl1, fl2, alm1, alm2, stream, add, pix1, pix2);
}
extern "C" int plk_map2alm_dev(plk_plan *p, int spin, const double *map1, const double *map2, const double *fl1,
                               const double *fl2, void *alm1, void *alm2, void *stream) {
  return map2alm_impl(p, spin, map1, map2, fl1, fl2, alm1, alm2, stream, AlmAdd());
}
// analysis with an additive per-l term folded into the output pass: alm_c = fl_c * analysis_c + afl_c[l] * add_c
// (afl_c: lmax + 1 doubles; add_c must not alias alm_c)
extern "C" int plk_map2alm_add_dev(plk_plan *p, int spin, const double *map1, const double *map2, const double *fl1,
                                   const double *fl2, const void *add1, const double *afl1, const void *add2,
                                   const double *afl2, void *alm1, void *alm2, void *stream) {
  if (!add1 || !afl1 || (spin > 0 && (!add2 || !afl2))) return fail(PLK_EINVAL, "NULL additive term");
  if (add1 == alm1 || (spin > 0 && add2 == alm2)) return fail(PLK_EINVAL, "additive term aliases the output");
  AlmAdd add;
  add.a1 = (const cplx *)add1; add.f1 = afl1; add.a2 = (const cplx *)add2; add.f2 = afl2;
  return map2alm_impl(p, spin, map1, map2, fl1, fl2, alm1, alm2, stream, add);
}

// ------------------------------------------------------------------------------------------ m-partitioned transforms
// Pure host arithmetic (callable without a GPU): ring-pair bounds per rank (balanced by pixel count) and the owner of
// every m column.
extern "C" int plk_dist_partition(int nside, int mmax, int nranks, int mblk, int *pair_lo, int *m_owner) {
  if (nside < 1 || (nside & (nside - 1)) || nranks < 1 || nranks > kMaxRanks || mblk < 1 || mmax < 0)
    return fail(PLK_EINVAL, "bad argument (nside %d, nranks %d, mblk %d)", nside, nranks, mblk);
  const int npair = 2 * nside;
  if (pair_lo) {
    // Cost of ring pair ip in the ring-FFT / pixel stage, in microseconds of one SM-parallel component pass, fitted to
    // the per-rank stage times of the nside-4096 / lmax-4000 'p' estimate on 8 B200 (profiles/r02_dist.md):
    //   power-of-two rings (q = 2^j, incl. the whole equatorial belt): 0.40 us at q = 4096, scaled with q log2 q
    //   Bluestein rings: 5.2e-6 * 2 M log2 M + 1.9e-5 q with M = nextpow2(2 q - 1) -- two M-point FFTs whatever q is,
    //   so the cost per PIXEL falls with q inside a size class (the round-1 model, pixels x a per-class factor, gave the
    //   polar ranks 7.1 ms of ring synthesis against 4.3 ms for the ranks holding the long Bluestein rings)
    std::vector<double> cum(npair + 1, 0.0);
    for (int ip = 0; ip < npair; ++ip) {
      const int q = ip < nside ? ip + 1 : nside;
      double c;
      if (q <= kTinyQ) c = 0.02 + 1.0e-5 * (mmax + 1);                       // direct sums over all m
      else if ((q & (q - 1)) == 0) c = 0.40 * (double)q * ilog2(q) / (4096.0 * 12.0);
      else { const int M = nextpow2(2 * q - 1); c = 5.2e-6 * 2.0 * M * ilog2(M) + 1.9e-5 * q; }
      if (ip == npair - 1) c *= 0.5;                                          // the equator has no southern twin
      cum[ip + 1] = cum[ip] + c;
    }
    pair_lo[0] = 0;
    for (int q = 1; q < nranks; ++q) {
      const double target = cum[npair] * q / nranks;
      int ip = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
      pair_lo[q] = std::min(std::max(ip, pair_lo[q - 1]), npair);
    }
    pair_lo[nranks] = npair;
  }
  if (m_owner) for (int m = 0; m <= mmax; ++m) m_owner[m] = dist_owner_of_m(m, mblk, nranks);
  return PLK_OK;
}

extern "C" int plk_dist_destroy(plk_dist *d) {
  if (!d) return PLK_OK;
  for (int q = 0; q < d->nranks; ++q)
    if (d->opened[q]) { cudaIpcCloseMemHandle(d->px1[q]); cudaIpcCloseMemHandle(d->px2[q]); }
  if (d->X1) cudaFree(d->X1);
  if (d->X2) cudaFree(d->X2);
  if (d->d_mlist) cudaFree(d->d_mlist);
  if (d->d_fft_list) cudaFree(d->d_fft_list);
  delete d;
  return PLK_OK;
}

extern "C" int plk_dist_create(plk_dist **out, plk_plan *p, int rank, int nranks, int mblk) {
  if (!out) return fail(PLK_EINVAL, "dist pointer is NULL");
  *out = nullptr;
  CHECK_PLAN(p);
  if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) return fail(PLK_EINVAL, "bad rank %d / %d", rank, nranks);
  if (mblk <= 0) mblk = 64;
  plk_dist *d = new plk_dist();
  d->plan = p; d->rank = rank; d->nranks = nranks; d->mblk = mblk;
  d->pair_lo.resize(nranks + 1);
  std::vector<int> owner(p->mmax + 1);
  int rc = plk_dist_partition(p->nside, p->mmax, nranks, mblk, d->pair_lo.data(), owner.data());
  if (rc) { delete d; return rc; }
  for (int m = 0; m <= p->mmax; ++m) if (owner[m] == rank) d->mlist.push_back(m);
#define DCK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { plk_dist_destroy(d); \
    return fail(PLK_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); } } while (0)
  DCK(cudaMalloc((void **)&d->d_mlist, std::max<size_t>(d->mlist.size(), 1) * sizeof(int)));
  if (!d->mlist.empty()) DCK(cudaMemcpy(d->d_mlist, d->mlist.data(), d->mlist.size() * sizeof(int), cudaMemcpyHostToDevice));
  // this rank's ring pairs, grouped by the plan's FFT size classes
  {
    std::vector<int> full(p->npair);
    DCK(cudaMemcpy(full.data(), p->fft_list, (size_t)p->npair * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int> list;
    for (const auto &c : p->fft_classes) {
      plk_plan::FftClass lc = c;
      lc.offset = (int)list.size();
      for (int i = 0; i < c.count; ++i) {
        const int ip = full[c.offset + i];
        if (ip >= d->pair_lo[rank] && ip < d->pair_lo[rank + 1]) list.push_back(ip);
      }
      lc.count = (int)list.size() - lc.offset;
      d->classes.push_back(lc);
    }
    DCK(cudaMalloc((void **)&d->d_fft_list, std::max<size_t>(list.size(), 1) * sizeof(int)));
    if (!list.empty()) DCK(cudaMemcpy(d->d_fft_list, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  d->tpitch = (p->nring + 1) & ~1;
  const size_t xb = std::max((size_t)p->nring * p->pitch, (size_t)(p->mmax + 1) * d->tpitch) * sizeof(cplx);
  DCK(cudaMalloc((void **)&d->X1, xb));
  DCK(cudaMalloc((void **)&d->X2, xb));
  DCK(cudaMemset(d->X1, 0, xb));
  DCK(cudaMemset(d->X2, 0, xb));
#undef DCK
  d->px1[rank] = d->X1; d->px2[rank] = d->X2;
  *out = d;
  return PLK_OK;
}

// CUDA IPC handles of the two phase arrays (2 x 64 bytes), to be exchanged between the processes of the box
extern "C" int plk_dist_export(plk_dist *d, void *handles) {
  if (!d || !handles) return fail(PLK_EINVAL, "NULL argument");
  cudaIpcMemHandle_t h[2];
  CK(cudaIpcGetMemHandle(&h[0], d->X1));
  CK(cudaIpcGetMemHandle(&h[1], d->X2));
  memcpy(handles, h, sizeof h);
  return PLK_OK;
}
extern "C" int plk_dist_import(plk_dist *d, int peer, const void *handles) {
  if (!d || !handles || peer < 0 || peer >= d->nranks) return fail(PLK_EINVAL, "bad argument");
  if (peer == d->rank) return PLK_OK;
  cudaIpcMemHandle_t h[2];
  memcpy(h, handles, sizeof h);
  void *a = nullptr, *b = nullptr;
  CK(cudaIpcOpenMemHandle(&a, h[0], cudaIpcMemLazyEnablePeerAccess));
  CK(cudaIpcOpenMemHandle(&b, h[1], cudaIpcMemLazyEnablePeerAccess));
  d->px1[peer] = (cplx *)a; d->px2[peer] = (cplx *)b; d->opened[peer] = true;
  return PLK_OK;
}
// same-process peers (several simulated ranks on one GPU, or one process driving peer-enabled GPUs)
extern "C" int plk_dist_set_peer(plk_dist *d, int peer, void *x1, void *x2) {
  if (!d || peer < 0 || peer >= d->nranks || !x1 || !x2) return fail(PLK_EINVAL, "bad argument");
  d->px1[peer] = (cplx *)x1; d->px2[peer] = (cplx *)x2;
  return PLK_OK;
}
extern "C" int plk_dist_phase_ptrs(plk_dist *d, void **x1, void **x2) {
  if (!d || !x1 || !x2) return fail(PLK_EINVAL, "NULL argument");
  *x1 = d->X1; *x2 = d->X2;
  return PLK_OK;
}
extern "C" int plk_dist_num_m(const plk_dist *d) { return d ? (int)d->mlist.size() : 0; }
// pixel index ranges [lo, hi) of this rank's rings in a RING map: north block, south block
extern "C" int plk_dist_pixel_ranges(const plk_dist *d, long long *r4) {
  if (!d || !r4) return fail(PLK_EINVAL, "NULL argument");
  const HostGeom &hg = d->plan->hg;
  const int lo = d->pair_lo[d->rank], hi = d->pair_lo[d->rank + 1];
  if (lo >= hi) { r4[0] = r4[1] = r4[2] = r4[3] = 0; return PLK_OK; }
  r4[0] = hg.start_n[lo];
  r4[1] = hg.start_n[hi - 1] + hg.nphi[hi - 1];
  // southern twins run backwards: pair hi-1 has the first southern ring of the block (the equator has none)
  int last = hi - 1;
  if (hg.start_s[last] < 0) --last;
  if (last < lo) { r4[2] = r4[3] = 0; return PLK_OK; }
  r4[2] = hg.start_s[last];
  r4[3] = hg.start_s[lo] + hg.nphi[lo];
  return PLK_OK;
}

static int dist_ready(const plk_dist *d) {
  if (!d) return fail(PLK_EINVAL, "dist is NULL");
  for (int q = 0; q < d->nranks; ++q)
    if (!d->px1[q] || !d->px2[q]) return fail(PLK_EINVAL, "peer %d has not been imported (plk_dist_import / plk_dist_set_peer)", q);
  return 0;
}
// Stage 1 of a distributed synthesis: Legendre sums of this rank's m columns over ALL ring pairs; every row is
// stored into the phase array of the rank that owns the ring pair (peer stores).  Callers place a barrier before
// (peers have consumed their arrays) and after (all rows have landed) -- see plancklens_b200/dist_sht.py.
extern "C" int plk_dist_legendre_synth(plk_dist *d, int spin, const void *alm1, const void *alm2, const double *fl1,
                                       const double *fl2, void *stream) {
  int rc = dist_ready(d);
  if (rc) return rc;
  if (spin < 0 || spin > 3) return fail(PLK_EINVAL, "spin must be 0..3, got %d", spin);
  if (!alm1) return fail(PLK_EINVAL, "NULL buffer");
  return legendre_synth(d->plan, spin, alm1, alm2, fl1, fl2, d->X1, d->X2, (cudaStream_t)stream, d);
}
// Stage 2: ring FFTs of this rank's ring pairs; writes only this rank's pixels of map1 (, map2)
extern "C" int plk_dist_ring_synth(plk_dist *d, int spin, double *map1, double *map2, void *stream) {
  int rc = dist_ready(d);
  if (rc) return rc;
  if (!map1 || (spin > 0 && !map2)) return fail(PLK_EINVAL, "NULL buffer");
  plk_plan *p = d->plan;
  if ((rc = ensure_spin(p, spin))) return rc;
  if ((rc = ensure_work(p))) return rc;
  const int *mtop = p->spins[spin].d.mtop;
  cudaStream_t st = (cudaStream_t)stream;
  // exchange buffer T[m][ring] -> phase array X[ring][m] for this rank's rings (north block, mirrored south block)
  const int lo = d->pair_lo[d->rank], hi = d->pair_lo[d->rank + 1];
  if (hi > lo) {
    const int rr[2][2] = {{lo, hi}, {p->nring - hi, p->nring - lo}};
    for (int c = 0; c < (spin ? 2 : 1); ++c)
      for (int k = 0; k < 2; ++k) {
        int r0 = rr[k][0], r1 = rr[k][1];
        if (k == 1 && r0 < hi) r0 = hi;          // the equator belongs to the north block
        if (r1 <= r0) continue;
        dim3 g((r1 - r0 + 31) / 32, (p->mmax + 32) / 32), b(32, 8);
        phase_transpose_kernel<<<g, b, 0, st>>>(c ? d->X2 : d->X1, d->tpitch, (cplx *)(c ? p->X2.p : p->X1.p), p->pitch, p->mmax, r0, r1);
        LAUNCHED();
      }
  }
  if ((rc = ring_synth(p, (const cplx *)p->X1.p, map1, st, mtop, d))) return rc;
  if (spin > 0 && (rc = ring_synth(p, (const cplx *)p->X2.p, map2, st, mtop, d))) return rc;
  return PLK_OK;
}
// Analysis stage 1: ring FFTs of this rank's rings; column m goes to the phase array of the rank owning m
extern "C" int plk_dist_ring_anal(plk_dist *d, int spin, const double *map1, const double *map2, void *stream) {
  int rc = dist_ready(d);
  if (rc) return rc;
  if (!map1 || (spin > 0 && !map2)) return fail(PLK_EINVAL, "NULL buffer");
  plk_plan *p = d->plan;
  if ((rc = ensure_spin(p, spin))) return rc;
  const int *mtop = p->spins[spin].d.mtop;
  if ((rc = ring_anal(p, map1, d->X1, (cudaStream_t)stream, mtop, d, 0))) return rc;
  if (spin > 0 && (rc = ring_anal(p, map2, d->X2, (cudaStream_t)stream, mtop, d, 1))) return rc;
  return PLK_OK;
}
// Analysis stage 2: Legendre analysis of this rank's m columns; rows of other ranks are set to zero
extern "C" int plk_dist_legendre_anal(plk_dist *d, int spin, const double *fl1, const double *fl2, void *alm1, void *alm2,
                                      void *stream) {
  int rc = dist_ready(d);
  if (rc) return rc;
  if (spin < 0 || spin > 3) return fail(PLK_EINVAL, "spin must be 0..3, got %d", spin);
  if (!alm1 || (spin > 0 && !alm2)) return fail(PLK_EINVAL, "NULL buffer");
  return legendre_anal(d->plan, spin, d->X1, d->X2, fl1, fl2, alm1, alm2, (cudaStream_t)stream, d);
}

// Same with the additive per-l term of plk_map2alm_add_dev applied to this rank's m rows only: the S^-1 x term of an
// m-distributed CG forward operator (rows of other ranks stay zero)
extern "C" int plk_dist_legendre_anal_add(plk_dist *d, int spin, const double *fl1, const double *fl2, const void *add1,
                                          const double *afl1, const void *add2, const double *afl2, void *alm1, void *alm2,
                                          void *stream) {
  int rc = dist_ready(d);
  if (rc) return rc;
  if (spin < 0 || spin > 3) return fail(PLK_EINVAL, "spin must be 0..3, got %d", spin);
  if (!alm1 || (spin > 0 && !alm2)) return fail(PLK_EINVAL, "NULL buffer");
  if (!add1 || !afl1 || (spin > 0 && (!add2 || !afl2))) return fail(PLK_EINVAL, "NULL additive term");
  if (add1 == alm1 || (spin > 0 && add2 == alm2)) return fail(PLK_EINVAL, "additive term aliases the output");
  AlmAdd add;
  add.a1 = (const cplx *)add1; add.f1 = afl1; add.a2 = (const cplx *)add2; add.f2 = afl2;
  return legendre_anal(d->plan, spin, d->X1, d->X2, fl1, fl2, alm1, alm2, (cudaStream_t)stream, d, add);
}

// ------------------------------------------------------------------------------------------ host-pointer variants
extern "C" int plk_alm2map_host(plk_plan *p, int spin, const void *alm1, const void *alm2, double *map1, double *map2) {
  CHECK_PLAN(p);
  if (spin < 0 || spin > 3) return fail(PLK_EINVAL, "spin must be 0..3, got %d", spin);
  if (!alm1 || !map1 || (spin > 0 && !map2)) return fail(PLK_EINVAL, "NULL buffer");
  const size_t ab = (size_t)alm_size(p->lmax, p->mmax) * sizeof(cplx), mb = (size_t)p->npix * sizeof(double);
  int rc;
  if ((rc = ensure(p->hostio[0], ab)) || (rc = ensure(p->hostio[2], mb))) return rc;
  CK(cudaMemcpy(p->hostio[0].p, alm1, ab, cudaMemcpyHostToDevice));
  void *a2 = nullptr;
  if (spin > 0) {
    if ((rc = ensure(p->hostio[3], mb))) return rc;
    if (alm2) {
      if ((rc = ensure(p->hostio[1], ab))) return rc;
      CK(cudaMemcpy(p->hostio[1].p, alm2, ab, cudaMemcpyHostToDevice));
      a2 = p->hostio[1].p;
    }
  }
  rc = plk_alm2map_dev(p, spin, p->hostio[0].p, a2, nullptr, nullptr, (double *)p->hostio[2].p, (double *)p->hostio[3].p, nullptr);
  if (rc) return rc;
  CK(cudaMemcpy(map1, p->hostio[2].p, mb, cudaMemcpyDeviceToHost));
  if (spin > 0) CK(cudaMemcpy(map2, p->hostio[3].p, mb, cudaMemcpyDeviceToHost));
  return PLK_OK;
}

extern "C" int plk_map2alm_host(plk_plan *p, int spin, const double *map1, const double *map2, void *alm1, void *alm2) {
  CHECK_PLAN(p);
  if (spin < 0 || spin > 3) return fail(PLK_EINVAL, "spin must be 0..3, got %d", spin);
  if (!alm1 || !map1 || (spin > 0 && (!map2 || !alm2))) return fail(PLK_EINVAL, "NULL buffer");
  const size_t ab = (size_t)alm_size(p->lmax, p->mmax) * sizeof(cplx), mb = (size_t)p->npix * sizeof(double);
  int rc;
  if ((rc = ensure(p->hostio[0], ab)) || (rc = ensure(p->hostio[2], mb))) return rc;
  CK(cudaMemcpy(p->hostio[2].p, map1, mb, cudaMemcpyHostToDevice));
  if (spin > 0) {
    if ((rc = ensure(p->hostio[1], ab)) || (rc = ensure(p->hostio[3], mb))) return rc;
    CK(cudaMemcpy(p->hostio[3].p, map2, mb, cudaMemcpyHostToDevice));
  }
  rc = plk_map2alm_dev(p, spin, (const double *)p->hostio[2].p, (const double *)p->hostio[3].p, nullptr, nullptr,
                       p->hostio[0].p, p->hostio[1].p, nullptr);
  if (rc) return rc;
  CK(cudaMemcpy(alm1, p->hostio[0].p, ab, cudaMemcpyDeviceToHost));
  if (spin > 0) CK(cudaMemcpy(alm2, p->hostio[1].p, ab, cudaMemcpyDeviceToHost));
  return PLK_OK;
}

// ------------------------------------------------------------------------------------------ BLAS-1 / pixel passes
static double *g_scratch = nullptr;   // per-process reduction scratch (current device at first use)
static unsigned int *g_ticket = nullptr;   // ticket counters of the last-block reductions (zero between launches)
static const size_t kScratchDoubles = 1 << 17;
// Lanes: work issued concurrently from several host threads on several streams (the T and the P filter of one simulation,
// filt_simple.library_sepTP.get_sim_teblm_dev) must not share reduction scratch.  A lane is a thread-local index chosen by
// the caller (plk_set_lane); each lane owns its slice of the scratch and its ticket counters, and launches recorded into a
// CUDA graph keep the slice of the lane they were captured in.
static thread_local int t_lane = 0;
static double *scr() { return g_scratch + (size_t)t_lane * kScratchDoubles; }
static unsigned int *tick() { return g_ticket + t_lane * 64; }
static std::mutex g_scratch_mu;
static int scratch() {
  std::lock_guard<std::mutex> lk(g_scratch_mu);
  if (g_scratch) return 0;
  double *s = nullptr;
  cudaError_t e = cudaMalloc((void **)&s, PLK_MAX_LANES * kScratchDoubles * sizeof(double));
  if (e != cudaSuccess) return fail(PLK_ENOMEM, "cudaMalloc(scratch) failed: %s", cudaGetErrorString(e));
  e = cudaMalloc((void **)&g_ticket, PLK_MAX_LANES * 64 * sizeof(unsigned int));
  if (e != cudaSuccess) return fail(PLK_ENOMEM, "cudaMalloc(ticket) failed: %s", cudaGetErrorString(e));
  e = cudaMemset(g_ticket, 0, PLK_MAX_LANES * 64 * sizeof(unsigned int));
  if (e != cudaSuccess) return fail(PLK_ECUDA, "cudaMemset(ticket) failed: %s", cudaGetErrorString(e));
  g_scratch = s;
  return 0;
}
extern "C" int plk_set_lane(int lane) {
  if (lane < 0 || lane >= PLK_MAX_LANES) return fail(PLK_EINVAL, "lane %d outside [0, %d)", lane, PLK_MAX_LANES);
  t_lane = lane;
  return PLK_OK;
}
extern "C" int plk_get_lane(void) { return t_lane; }
static int flat_grid(long long n) { return (int)std::min<long long>((n + 255) / 256, 148 * 16); }

extern "C" int plk_almxfl_dev(int lmax, const void *in, const double *fl, int nfl, void *out, void *stream) {
  if (!in || !out || !fl || lmax < 0) return fail(PLK_EINVAL, "bad argument");
  dim3 g((lmax + 256) / 256, lmax + 1);
  almxfl_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(lmax, (const cplx *)in, fl, nfl, (cplx *)out);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_alm_axpy_dev(long long n, double a, const double *a_dev, const void *x, void *y, void *stream) {
  if (!x || !y || n < 0) return fail(PLK_EINVAL, "bad argument");
  axpy_kernel<<<flat_grid(2 * n), 256, 0, (cudaStream_t)stream>>>(2 * n, a, a_dev, (const double *)x, (double *)y);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_alm_dot_dev(int lmax, int lmin, const void *a, const void *b, double *result_dev, void *stream) {
  if (!a || !b || !result_dev || lmax < 0) return fail(PLK_EINVAL, "bad argument");
  int rc = scratch();
  if (rc) return rc;
  if ((size_t)lmax + 1 > kScratchDoubles) return fail(PLK_EINVAL, "lmax too large");
  dot_partial_kernel<<<lmax + 1, 256, 0, (cudaStream_t)stream>>>(lmax, lmin, (const cplx *)a, (const cplx *)b, scr());
  LAUNCHED();
  final_sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(scr(), lmax + 1, 1, result_dev);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_alm_dot2_dev(int lmax, int lmin, const void *a1, const void *b1, const void *a2, const void *b2,
                                double *result_dev, void *stream) {
  if (!a1 || !b1 || !a2 || !b2 || !result_dev || lmax < 0) return fail(PLK_EINVAL, "bad argument");
  int rc = scratch();
  if (rc) return rc;
  if (2 * ((size_t)lmax + 1) > kScratchDoubles) return fail(PLK_EINVAL, "lmax too large");
  dot_partial_kernel<<<lmax + 1, 256, 0, (cudaStream_t)stream>>>(lmax, lmin, (const cplx *)a1, (const cplx *)b1, scr());
  LAUNCHED();
  dot_partial_kernel<<<lmax + 1, 256, 0, (cudaStream_t)stream>>>(lmax, lmin, (const cplx *)a2, (const cplx *)b2, scr() + lmax + 1);
  LAUNCHED();
  final_sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(scr(), 2 * (lmax + 1), 1, result_dev);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_alm_dotn_dev(int lmax, int lmin, int n, const void *const *a, const void *const *b, double *result_dev,
                                void *stream) {
  if (!a || !b || !result_dev || lmax < 0 || n < 1 || n > 4) return fail(PLK_EINVAL, "bad argument");
  int rc = scratch();
  if (rc) return rc;
  if ((size_t)n * ((size_t)lmax + 1) > kScratchDoubles) return fail(PLK_EINVAL, "lmax too large");
  for (int j = 0; j < n; ++j) {
    if (!a[j] || !b[j]) return fail(PLK_EINVAL, "NULL component");
    dot_partial_kernel<<<lmax + 1, 256, 0, (cudaStream_t)stream>>>(lmax, lmin, (const cplx *)a[j], (const cplx *)b[j],
                                                                    scr() + (size_t)j * (lmax + 1));
    LAUNCHED();
  }
  final_sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(scr(), n * (lmax + 1), 1, result_dev);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_alm_dot_fused_dev(int lmax, int lmin, int n, const void *const *a, const void *const *b, const double *num,
                                     const double *den, double scale, double *out3, void *stream) {
  if (!a || !b || !out3 || lmax < 0 || n < 1 || n > 4) return fail(PLK_EINVAL, "bad argument");
  if (num && den) return fail(PLK_EINVAL, "give num or den, not both");
  int rc = scratch();
  if (rc) return rc;
  if ((size_t)n * ((size_t)lmax + 1) > kScratchDoubles / 2) return fail(PLK_EINVAL, "lmax too large");
  DotArgs q;
  q.n = n;
  for (int j = 0; j < 4; ++j) { q.a[j] = nullptr; q.b[j] = nullptr; }
  for (int j = 0; j < n; ++j) {
    if (!a[j] || !b[j]) return fail(PLK_EINVAL, "NULL component");
    q.a[j] = (const cplx *)a[j]; q.b[j] = (const cplx *)b[j];
  }
  // second half of the scratch: the first half belongs to the two-kernel dots, which may be in flight on the stream
  dot_fused_kernel<<<n * (lmax + 1), 256, 0, (cudaStream_t)stream>>>(q, lmax, lmin, scr() + kScratchDoubles / 2, tick(),
                                                                      num, den, scale, out3);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_alm_axpy2_dev(long long n, const double *a_dev, const void *x1, void *y1, const void *x2, void *y2, void *stream) {
  if (!a_dev || !x1 || !y1 || !x2 || !y2 || n < 0) return fail(PLK_EINVAL, "bad argument");
  axpy2_kernel<<<flat_grid(2 * n), 256, 0, (cudaStream_t)stream>>>(2 * n, a_dev, (const double *)x1, (double *)y1,
                                                                    (const double *)x2, (double *)y2);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_alm2cl_dev(int lmax, const void *a, const void *b, double *cl, void *stream) {
  if (!a || !b || !cl || lmax < 0) return fail(PLK_EINVAL, "bad argument");
  alm2cl_kernel<<<(lmax + 128) / 128, 128, 0, (cudaStream_t)stream>>>(lmax, (const cplx *)a, (const cplx *)b, cl);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_scalar_ratio_dev(const double *num, const double *den, double scale, double *out, void *stream) {
  if (!num || !den || !out) return fail(PLK_EINVAL, "NULL buffer");
  scalar_ratio_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(num, den, scale, out);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_alm_copy_dev(int lmax_in, const void *in, int lmax_out, void *out, void *stream) {
  if (!in || !out) return fail(PLK_EINVAL, "NULL buffer");
  dim3 g((lmax_out + 256) / 256, lmax_out + 1);
  alm_copy_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(lmax_in, (const cplx *)in, lmax_out, (cplx *)out);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_alm_splice_dev(int lmax_lo, const void *lo, int lmax_hi, const void *hi, int lsplit, void *out,
                                  void *stream) {
  if (!lo || !hi || !out) return fail(PLK_EINVAL, "NULL buffer");
  if (lsplit > lmax_lo || lsplit > lmax_hi) return fail(PLK_EINVAL, "lsplit exceeds lmax");
  dim3 g((lmax_hi + 256) / 256, lmax_hi + 1);
  alm_splice_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(lmax_lo, (const cplx *)lo, lmax_hi, (const cplx *)hi, lsplit, (cplx *)out);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_map_mul_dev(long long n, double *y, const double *a, void *stream) {
  if (!y || !a) return fail(PLK_EINVAL, "NULL buffer");
  map_mul_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(n, y, a);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_map_dot_dev(long long n, const double *a, const double *b, double *result_dev, void *stream) {
  if (!a || !b || !result_dev || n < 0) return fail(PLK_EINVAL, "bad argument");
  int rc = scratch();
  if (rc) return rc;
  const int nb = 148 * 8;
  map_dot_partial_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(n, a, b, scr());
  LAUNCHED();
  final_sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(scr(), nb, 1, result_dev);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_map_mul2_dev(long long n, double *g, double *c, const double *t, void *stream) {
  if (!g || !c || !t) return fail(PLK_EINVAL, "NULL buffer");
  map_mul2_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(n, g, c, t);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_map_qe_pp_dev(long long n, const double *q, const double *u, const double *g3, const double *c3,
                                 const double *g1, const double *c1, double *re, double *im, void *stream) {
  if (!q || !u || !g3 || !c3 || !g1 || !c1 || !re || !im) return fail(PLK_EINVAL, "NULL buffer");
  map_qe_pp_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(n, q, u, g3, c3, g1, c1, re, im);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_map_cmul_acc_dev(long long n, const double *ar, const double *ai, const double *br, const double *bi,
                                    double *dr, double *di, void *stream) {
  if (!ar || !br || !dr || !di) return fail(PLK_EINVAL, "NULL buffer");
  map_cmul_acc_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(n, ar, ai, br, bi, dr, di);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_map_ninv3_dev(long long n, double *q, double *u, const double *nqq, const double *nqu,
                                 const double *nuu, void *stream) {
  if (!q || !u || !nqq || !nqu || !nuu) return fail(PLK_EINVAL, "NULL buffer");
  map_ninv3_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(n, q, u, nqq, nqu, nuu);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_udgrade_sum_dev(int nside_in, const double *in, int nside_out, double *out, void *stream) {
  if (!in || !out) return fail(PLK_EINVAL, "NULL buffer");
  if (nside_in < 1 || nside_out < 1 || (nside_in & (nside_in - 1)) || (nside_out & (nside_out - 1)) || nside_out > nside_in ||
      nside_in > 8192)
    return fail(PLK_EINVAL, "nside_in (%d) and nside_out (%d) must be powers of two, nside_out <= nside_in <= 8192", nside_in, nside_out);
  udgrade_sum_kernel<<<flat_grid(12LL * nside_out * nside_out), 256, 0, (cudaStream_t)stream>>>(nside_in, in, nside_out, out);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_map_modes_dot_dev(plk_plan *p, double *m, const double *w, double *sums_dev, void *stream) {
  CHECK_PLAN(p);
  if (!m || !sums_dev) return fail(PLK_EINVAL, "NULL buffer");
  int rc = ensure(p->partial, (size_t)4 * p->nring * sizeof(double));
  if (rc) return rc;
  if ((rc = scratch())) return rc;
  modes_dot_kernel<<<p->nring, 256, 0, (cudaStream_t)stream>>>(p->rings, m, w, (double *)p->partial.p, tick() + 1, sums_dev);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_map_modes_sub_dev(plk_plan *p, double *m, const double *w, const double *sums_dev,
                                     const double *pinv_dev, void *stream) {
  CHECK_PLAN(p);
  if (!m || !sums_dev || !pinv_dev) return fail(PLK_EINVAL, "NULL buffer");
  modes_sub_kernel<<<p->nring, 256, 0, (cudaStream_t)stream>>>(p->rings, m, w, sums_dev, pinv_dev);
  LAUNCHED();
  return PLK_OK;
}

extern "C" int plk_alm_lincomb_dev(long long n, double ca, const void *x, double cb, const void *y, void *out, void *stream) {
  if (!x || !out || n < 0) return fail(PLK_EINVAL, "bad argument");
  lincomb_kernel<<<flat_grid(2 * n), 256, 0, (cudaStream_t)stream>>>(2 * n, ca, (const double *)x, cb, (const double *)y, (double *)out);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_alm_combine_dev(int lmax, int nterm, const void *const *in, const double *const *fl, const int *nfl,
                                   void *out, void *stream) {
  if (nterm < 1 || nterm > 4 || !in || !fl || !nfl || !out) return fail(PLK_EINVAL, "bad argument");
  AlmTerms t;
  t.nterm = nterm;
  for (int j = 0; j < 4; ++j) { t.in[j] = nullptr; t.fl[j] = nullptr; t.nfl[j] = 0; }
  for (int j = 0; j < nterm; ++j) {
    if (!in[j] || !fl[j]) return fail(PLK_EINVAL, "NULL term");
    t.in[j] = (const cplx *)in[j]; t.fl[j] = fl[j]; t.nfl[j] = nfl[j];
  }
  dim3 g((lmax + 256) / 256, lmax + 1);
  alm_combine_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(lmax, t, (cplx *)out);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_alm2rlm_dev(int lmax, const void *alm, double *rlm, void *stream) {
  if (!alm || !rlm) return fail(PLK_EINVAL, "NULL buffer");
  dim3 g((lmax + 256) / 256, lmax + 1);
  alm2rlm_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(lmax, (const cplx *)alm, rlm, lmax);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_alm2rlm_from_dev(int lmax, int lmax_src, const void *alm, double *rlm, void *stream) {
  if (!alm || !rlm) return fail(PLK_EINVAL, "NULL buffer");
  if (lmax_src < lmax) return fail(PLK_EINVAL, "source lmax (%d) below the packed lmax (%d)", lmax_src, lmax);
  dim3 g((lmax + 256) / 256, lmax + 1);
  alm2rlm_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(lmax, (const cplx *)alm, rlm, lmax_src);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_alm_splice_xfl_dev(int lmax_lo, const void *lo, int lmax_hi, const void *hi, const double *fl, int nfl,
                                      int lsplit, void *out, void *stream) {
  if (!lo || !hi || !fl || !out) return fail(PLK_EINVAL, "NULL buffer");
  if (lsplit > lmax_lo || lsplit > lmax_hi) return fail(PLK_EINVAL, "lsplit exceeds lmax");
  dim3 g((lmax_hi + 256) / 256, lmax_hi + 1);
  alm_splice_xfl_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(lmax_lo, (const cplx *)lo, lmax_hi, (const cplx *)hi, fl, nfl, lsplit, (cplx *)out);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_rlm2alm_dev(int lmax, const double *rlm, void *alm, void *stream) {
  if (!alm || !rlm) return fail(PLK_EINVAL, "NULL buffer");
  dim3 g((lmax + 256) / 256, lmax + 1);
  rlm2alm_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(lmax, rlm, (cplx *)alm);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_dense_matvec_dev(int n, const double *A, const double *x, double *y, void *stream) {
  if (!A || !x || !y || n < 1) return fail(PLK_EINVAL, "bad argument");
  matvec_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(n, A, x, y);
  LAUNCHED();
  return PLK_OK;
}

// ------------------------------------------------------------------------------------------ device random numbers
extern "C" int plk_randn_dev(unsigned long long seed, unsigned long long stream_id, long long n, double scale,
                             const double *add, double *out, void *stream) {
  if (!out || n < 0) return fail(PLK_EINVAL, "bad argument");
  if (((uintptr_t)out & 15) || (add && ((uintptr_t)add & 15))) return fail(PLK_EINVAL, "buffers must be 16-byte aligned");
  if (n == 0) return PLK_OK;
  randn_kernel<<<flat_grid((n + 1) / 2), 256, 0, (cudaStream_t)stream>>>(seed, stream_id, n, scale, add, out);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_randn_alm_dev(unsigned long long seed, unsigned long long stream_id, int lmax, void *alm, void *stream) {
  if (!alm || lmax < 0) return fail(PLK_EINVAL, "bad argument");
  const long long nalm = alm_size(lmax, lmax);
  randn_alm_kernel<<<flat_grid(nalm), 256, 0, (cudaStream_t)stream>>>(seed, stream_id, nalm, lmax, (cplx *)alm);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_philox_words_dev(unsigned long long seed, unsigned long long stream_id, long long ncalls,
                                    unsigned int *out, void *stream) {
  if (!out || ncalls < 0) return fail(PLK_EINVAL, "bad argument");
  if (ncalls == 0) return PLK_OK;
  philox_words_kernel<<<flat_grid(ncalls), 256, 0, (cudaStream_t)stream>>>(seed, stream_id, ncalls, out);
  LAUNCHED();
  return PLK_OK;
}

// ------------------------------------------------------------------------------------------ Wigner small-d transforms
namespace {
struct WigDev {
  double *A = nullptr, *B = nullptr, *C = nullptr;
  WigCoef w{};
};
std::map<std::tuple<int, int, int>, WigDev> g_wig;     // (s1, s2, lmax) -> device coefficient tables (current device)
DevBuf g_wig_partial, g_wig_clw;

int wig_get(int s1, int s2, int lmax, WigCoef *out) {
  const auto key = std::make_tuple(s1, s2, lmax);
  auto it = g_wig.find(key);
  if (it == g_wig.end()) {
    const int l0 = std::max(std::abs(s1), std::abs(s2));
    std::vector<double> A(lmax + 2, 0.0), B(lmax + 2, 0.0), C(lmax + 2, 0.0);
    const long double m = s1, mp = s2;
    for (int l = l0; l <= lmax; ++l) {
      if (l == 0) { A[l] = 1.0; continue; }             // P_1 = x P_0
      const long double j = l;
      const long double den = j * sqrtl(((j + 1) * (j + 1) - m * m) * ((j + 1) * (j + 1) - mp * mp));
      A[l] = (double)((2 * j + 1) * j * (j + 1) / den);
      B[l] = (double)(-(2 * j + 1) * m * mp / den);
      C[l] = (double)((j + 1) * sqrtl((j * j - m * m) * (j * j - mp * mp)) / den);
    }
    WigDev d;
    const size_t nb = (size_t)(lmax + 2) * sizeof(double);
    CK(cudaMalloc((void **)&d.A, nb)); CK(cudaMalloc((void **)&d.B, nb)); CK(cudaMalloc((void **)&d.C, nb));
    CK(cudaMemcpy(d.A, A.data(), nb, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.B, B.data(), nb, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.C, C.data(), nb, cudaMemcpyHostToDevice));
    const int a = std::abs(s1 - s2), b = std::abs(s1 + s2);
    long double seed = 1.0L;                             // sqrt((a+b)! / (a! b!)) = sqrt(binomial(a+b, a))
    for (int k = 1; k <= a; ++k) seed *= (long double)(b + k) / k;
    seed = sqrtl(seed);
    if (s2 < s1 && ((s1 - s2) & 1)) seed = -seed;        // xi_{m m'} = (-1)^{m' - m} for m' < m
    d.w.A = d.A; d.w.B = d.B; d.w.C = d.C; d.w.l0 = l0; d.w.lmax = lmax; d.w.seed = (double)seed; d.w.a = a; d.w.b = b;
    it = g_wig.emplace(key, d).first;
  }
  *out = it->second.w;
  return 0;
}
}  // namespace

extern "C" int plk_wignerpos_dev(const double *cl, int lmax, const double *x, int nx, int s1, int s2, double *xi, void *stream) {
  if (!cl || !x || !xi || lmax < 0 || nx < 1) return fail(PLK_EINVAL, "bad argument");
  if (std::abs(s1) > 8 || std::abs(s2) > 8) return fail(PLK_EINVAL, "spins out of range: %d %d", s1, s2);
  cudaStream_t st = (cudaStream_t)stream;
  if (std::max(std::abs(s1), std::abs(s2)) > lmax) { CK(cudaMemsetAsync(xi, 0, (size_t)nx * sizeof(double), st)); return PLK_OK; }
  WigCoef w;
  int rc = wig_get(s1, s2, lmax, &w);
  if (rc) return rc;
  if ((rc = ensure(g_wig_clw, (size_t)(lmax + 1) * sizeof(double)))) return rc;
  wig_scale_kernel<<<(lmax + 256) / 256, 256, 0, st>>>(cl, lmax, (double *)g_wig_clw.p);
  LAUNCHED();
  wignerpos_kernel<<<(nx + 127) / 128, 128, 0, st>>>(w, (const double *)g_wig_clw.p, x, nx, xi);
  LAUNCHED();
  return PLK_OK;
}
extern "C" int plk_wignercoeff_dev(const double *f, const double *x, int nx, int s1, int s2, int lmax, double *cl, void *stream) {
  if (!cl || !x || !f || lmax < 0 || nx < 1) return fail(PLK_EINVAL, "bad argument");
  if (std::abs(s1) > 8 || std::abs(s2) > 8) return fail(PLK_EINVAL, "spins out of range: %d %d", s1, s2);
  cudaStream_t st = (cudaStream_t)stream;
  if (std::max(std::abs(s1), std::abs(s2)) > lmax) { CK(cudaMemsetAsync(cl, 0, (size_t)(lmax + 1) * sizeof(double), st)); return PLK_OK; }
  WigCoef w;
  int rc = wig_get(s1, s2, lmax, &w);
  if (rc) return rc;
  const int nblk = (nx + 255) / 256, pitch = lmax + 1;
  if ((rc = ensure(g_wig_partial, (size_t)nblk * pitch * sizeof(double)))) return rc;
  wignercoeff_kernel<<<nblk, 256, 0, st>>>(w, f, x, nx, (double *)g_wig_partial.p, pitch);
  LAUNCHED();
  wigner_finish_kernel<<<(lmax + 256) / 256, 256, 0, st>>>((const double *)g_wig_partial.p, nblk, pitch, lmax, 2.0 * M_PI, cl);
  LAUNCHED();
  return PLK_OK;
}

// ------------------------------------------------------------------------------------------ measurement helpers
extern "C" int plk_profile_enable(int on) {
  for (auto &r : g_prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof_recs.clear();
  g_prof = on != 0;
  return PLK_OK;
}
// sums the recorded durations per kind (5 entries each): counts[k], total_ms[k]; synchronises the device
extern "C" int plk_profile_read(int *counts, double *total_ms) {
  CK(cudaDeviceSynchronize());
  for (int k = 0; k < 5; ++k) { counts[k] = 0; total_ms[k] = 0.0; }
  for (auto &r : g_prof_recs) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, r.e0, r.e1));
    counts[r.kind] += 1; total_ms[r.kind] += ms;
  }
  return PLK_OK;
}

// Share of the (l, m, ring pair) volume the Legendre kernels actually walk for this spin: 1 - the part below the
// 2^-60 start threshold near the poles (skipped).  bench.py reports the roofline both on the algorithmic count of
// SURVEY.md section 8d (full volume) and on this executed share.
extern "C" int plk_plan_active_fraction(plk_plan *p, int spin, double *frac) {
  CHECK_PLAN(p);
  if (!frac) return fail(PLK_EINVAL, "NULL");
  int rc = ensure_spin(p, spin);
  if (rc) return rc;
  const size_t ns = (size_t)(p->mmax + 1) * p->npair;
  std::vector<int> ks(ns);
  CK(cudaMemcpy(ks.data(), p->spins[spin].d.ks, ns * sizeof(int), cudaMemcpyDeviceToHost));
  double act = 0.0, tot = 0.0;
  for (int m = 0; m <= p->mmax; ++m) {
    const int l0 = m > spin ? m : spin;
    const int K = p->lmax - l0 + 1;
    if (K <= 0) continue;
    for (int ip = 0; ip < p->npair; ++ip) {
      const int k = ks[(size_t)m * p->npair + ip];
      tot += K;
      if (k < K) act += K - k;
    }
  }
  *frac = tot > 0 ? act / tot : 0.0;
  return PLK_OK;
}

// FP64 FMA peak of the device: 8 independent dependent-chains of DFMA per thread, 1024 threads per SM resident
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b) {
  double v0 = threadIdx.x * 1e-3, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
      v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
}
// returns the measured rate in TFLOP/s (2 flop per FMA), best of `reps`
extern "C" int plk_fp64_peak(double *tflops, int reps) {
  if (!tflops) return fail(PLK_EINVAL, "NULL");
  int dev; CK(cudaGetDevice(&dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  double *d; CK(cudaMalloc((void **)&d, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int r = 0; r < reps + 1; ++r) {
    cudaEventRecord(e0);
    fp64_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    g_launches.fetch_add(1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 64.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  *tflops = best;
  return PLK_OK;
}
