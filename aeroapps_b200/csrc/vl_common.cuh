// Stage B (VLIDORT) -- shared definitions for the CUDA kernels.
//
// The numerical bodies are written against a small "Team" abstraction (a group of cooperating lanes: a warp or
// a whole CTA).  Every loop has the form `for (i = T.rank; i < count; i += T.size)` followed by T.sync(), and
// cross-lane reductions go through T.sum()/T.argmax().  On the GPU a Team is a warp (WarpTeam) or a thread
// block (BlockTeam); tests/hostmodel builds the very same bodies with a one-lane SerialTeam so that the
// mathematics can be checked against the CPU oracle in the GPU-less build container (test-only build: the
// shipped library contains the CUDA instantiations only and has no CPU path).
#pragma once
#include <cmath>
#include <cstdint>

#ifdef AB_HOST_MODEL
#define DEVFN inline
#define HDFN inline
#define AB_RESTRICT
#else
#include <cuda_runtime.h>
#define DEVFN __device__ __forceinline__
#define HDFN __host__ __device__ inline
#define AB_RESTRICT __restrict__
#endif

namespace vlk {

// vlidort_pars.f90:380-440
constexpr double DEG_TO_RAD = 1.7453292519943e-02;     // truncated on purpose (reference value)
constexpr double PIE = 180.0 * DEG_TO_RAD;
constexpr double PI4 = 4.0 * PIE;
constexpr double TAYLOR_SMALL = 1.0e-4;
constexpr double OMEGA_SMALLNUM = 1.0e-15;
constexpr double MAX_TAU_PATH = 32.0;                  // MAX_TAU_SPATH = _UPATH = _QPATH
constexpr int TAYLOR_ORDER = 3;
constexpr int MAXMOM = 64;                             // >= 2*nstreams
constexpr int MAXLAY = 128;

#ifdef AB_HOST_MODEL
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }     // CUDA's rsqrt on the host
#endif

struct SerialTeam {
  static constexpr int rank = 0, size = 1;
  static constexpr bool is_cta = false;       // true: a whole CTA (threadIdx/blockDim/__syncthreads forms of vl_linalg.cuh apply)
  DEVFN void sync() const {}
  DEVFN double sum(double v) const { return v; }
  DEVFN double bcast0(double v) const { return v; }
  // argmax over lanes: returns max value, idx receives the owner's idx
  DEVFN double argmax(double v, int &idx) const { return v; }
  // "panel" sub-team (first PW ranks): warp-scope collectives without a block barrier
  static constexpr int PW = 1, PRU = 256;                  // lanes, max rows per lane
  DEVFN double pargmax(double v, int &idx) const { return v; }
  template <int N_>
  DEVFN void pbcast_n(double (&)[N_], int, int) const {}
  DEVFN void psync() const {}
  // bulk load of n doubles into team scratch: begin (collective) ... wait (collective)
  DEVFN void load_begin(double *dst, const double *src, size_t n) const { for (size_t i = 0; i < n; i++) dst[i] = src[i]; }
  DEVFN void load_wait() const {}
  // next item of a shared work counter, one value per 32-lane group of the team
  DEVFN int wfetch(int *ctr) const { return (*ctr)++; }
  // asynchronous copy of n doubles global -> team scratch on one of 4 channels: begin (collective) ... wait (collective)
  DEVFN void copy_begin(int, double *dst, const double *src, size_t n) const { for (size_t i = 0; i < n; i++) dst[i] = src[i]; }
  DEVFN void copy_wait(int) const {}
};

#ifndef AB_HOST_MODEL
DEVFN unsigned vl_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
struct WarpTeam {
  int rank;
  static constexpr int size = 32;
  static constexpr bool is_cta = false;
  DEVFN WarpTeam() : rank(threadIdx.x & 31) {}
  DEVFN void sync() const { __syncwarp(); }
  DEVFN double sum(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  DEVFN double bcast0(double v) const { return __shfl_sync(0xffffffffu, v, 0); }
  DEVFN double argmax(double v, int &idx) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, v, o);
      int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    return v;
  }
  // the K3 interface (panel sub-team = the whole warp): lets one warp run a whole boundary-value problem when nstokes*nstreams is
  // small (3n <= 64 window rows), several pixels per CTA
  static constexpr int PW = 32, PRU = 2;
  DEVFN void psync() const { __syncwarp(); }
  DEVFN double pargmax(double v, int &idx) const {
    const unsigned long long key = v < 0.0 ? 0ull : (unsigned long long)__double_as_longlong(v);
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    const bool cand = (hi == mh) && (lo == ml);
    idx = (int)__reduce_min_sync(0xffffffffu, cand ? (unsigned)idx : 0x7fffffffu);
    return __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
  }
  template <int N_>
  DEVFN void pbcast_n(double (&v)[N_], int j0, int owner) const {
#pragma unroll
    for (int j = 0; j < N_; j++) if (j >= j0) v[j] = __shfl_sync(0xffffffffu, v[j], owner);
  }
  DEVFN int wfetch(int *ctr) const {
    int g = 0;
    if (rank == 0) g = atomicAdd(ctr, 1);
    return __shfl_sync(0xffffffffu, g, 0);
  }
  DEVFN void load_begin(double *dst, const double *src, size_t n) const {
    __syncwarp();
    for (size_t i = rank; i < n; i += 32) dst[i] = src[i];
    __syncwarp();
  }
  DEVFN void load_wait() const {}
  DEVFN void copy_begin(int, double *dst, const double *src, size_t n) const { load_begin(dst, src, n); }
  DEVFN void copy_wait(int) const {}
};

// W = 8 or 16 neighbouring lanes of a warp working on one problem (small nstokes*nstreams: several problems per warp); every
// collective carries the sub-team's lane mask, so the sub-teams of a warp may diverge freely
template <int W>
struct SubWarpTeam {
  int rank;
  static constexpr int size = W;
  unsigned mask;
  DEVFN SubWarpTeam() : rank(threadIdx.x & (W - 1)), mask(((1u << W) - 1u) << ((threadIdx.x & 31) & ~(W - 1))) {}
  DEVFN void sync() const { __syncwarp(mask); }
  DEVFN double sum(double v) const {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
  }
  DEVFN double bcast0(double v) const { return __shfl_sync(mask, v, 0, W); }
  DEVFN double argmax(double v, int &idx) const {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(mask, v, o);
      int oi = __shfl_xor_sync(mask, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    return v;
  }
};

// two warps (2j, 2j+1 of the CTA, i.e. two different schedulers) working on one problem: K2 with twice the warps per SM for the
// same shared-memory footprint.  Collectives: named barrier 1 + j (64 threads) and a 4-double exchange area `xch`.
template <int MAXHALF>
struct DuoTeamT {
  int rank;
  static constexpr int size = 64;
  static constexpr int maxhalf = MAXHALF;       // rows of a Jacobi half-pair kept in registers (18: n <= 36, 24: n <= 48)
  int bar;
  double *xch;
  DEVFN DuoTeamT(double *exchange) : rank(((threadIdx.x >> 5) & 1) * 32 + (threadIdx.x & 31)), bar(1 + (threadIdx.x >> 6)), xch(exchange) {}
  DEVFN void sync() const { asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory"); }
  DEVFN double sum(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    sync();
    if ((rank & 31) == 0) xch[rank >> 5] = v;
    sync();
    return xch[0] + xch[1];
  }
  DEVFN double bcast0(double v) const { sync(); if (rank == 0) xch[2] = v; sync(); return xch[2]; }
  DEVFN double argmax(double v, int &idx) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, v, o);
      int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    sync();
    if ((rank & 31) == 0) { xch[rank >> 5] = v; xch[2 + (rank >> 5)] = (double)idx; }
    sync();
    const double v0 = xch[0], v1 = xch[1]; const int i0 = (int)xch[2], i1 = (int)xch[3];
    if (v1 > v0 || (v1 == v0 && i1 < i0)) { idx = i1; return v1; }
    idx = i0; return v0;
  }
};
using DuoTeam = DuoTeamT<18>;

// whole CTA; scratch must hold 64 doubles + 64 ints
struct BlockTeam {
  int rank, size;
  static constexpr bool is_cta = true;
  double *sd; int *si;
  unsigned long long *mbar;      // optional mbarrier (shared memory) for load_begin/load_wait
  mutable unsigned mphase;
  unsigned long long *cbar;      // optional 4 mbarriers for the copy_begin/copy_wait channels
  mutable unsigned cphase;       // bit c = phase parity of channel c
  DEVFN BlockTeam(double *scratch_d, int *scratch_i, unsigned long long *bar = nullptr, unsigned long long *chan = nullptr)
      : rank(threadIdx.x), size(blockDim.x), sd(scratch_d), si(scratch_i), mbar(bar), mphase(0), cbar(chan), cphase(0) {
    if (bar || chan) {
      if (threadIdx.x == 0) {
        if (bar) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vl_smem_u32(bar)), "r"(1));
        if (chan) for (int c = 0; c < 4; c++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vl_smem_u32(chan + c)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      __syncthreads();
    }
  }
  // channel c: 1-D TMA bulk copy global -> shared of n doubles (n*8 a multiple of 16, both sides 16-byte aligned), completion on the
  // channel's mbarrier.  Collective: the barrier inside orders the team's earlier accesses to dst before the asynchronous write.
  DEVFN void copy_begin(int c, double *dst, const double *src, size_t n) const {
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    if (rank == 0) {
      const unsigned bytes = (unsigned)(n * 8);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(vl_smem_u32(cbar + c)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(vl_smem_u32(dst)),
                   "l"(src), "r"(bytes), "r"(vl_smem_u32(cbar + c))
                   : "memory");
    }
  }
  DEVFN void copy_wait(int c) const {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "VL_CWAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra VL_CDONE;\n"
        "bra VL_CWAIT;\n"
        "VL_CDONE:\n"
        "}" ::"r"(vl_smem_u32(cbar + c)),
        "r"((cphase >> c) & 1u)
        : "memory");
    cphase ^= 1u << c;
  }
  // 1-D TMA bulk copy global -> shared (cp.async.bulk), n*8 bytes a multiple of 16, both sides 16-byte aligned.
  // Collective: orders the team's earlier generic-proxy accesses to dst/src before the asynchronous copy.
  DEVFN void load_begin(double *dst, const double *src, size_t n) const {
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    if (rank == 0) {
      const unsigned bytes = (unsigned)(n * 8);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(vl_smem_u32(mbar)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(vl_smem_u32(dst)),
                   "l"(src), "r"(bytes), "r"(vl_smem_u32(mbar))
                   : "memory");
    }
  }
  DEVFN void load_wait() const {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "VL_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra VL_DONE;\n"
        "bra VL_WAIT;\n"
        "VL_DONE:\n"
        "}" ::"r"(vl_smem_u32(mbar)),
        "r"(mphase)
        : "memory");
    mphase ^= 1u;
  }
  DEVFN void sync() const { __syncthreads(); }
  DEVFN double sum(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((rank & 31) == 0) sd[rank >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (size + 31) / 32; w++) t += sd[w];
    return t;
  }
  DEVFN double bcast0(double v) const {
    __syncthreads();
    if (rank == 0) sd[0] = v;
    __syncthreads();
    return sd[0];
  }
  DEVFN double argmax(double v, int &idx) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, v, o);
      int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    __syncthreads();
    if ((rank & 31) == 0) { sd[rank >> 5] = v; si[rank >> 5] = idx; }
    __syncthreads();
    double bv = sd[0]; int bi = si[0];
    for (int w = 1; w < (size + 31) / 32; w++) { if (sd[w] > bv || (sd[w] == bv && si[w] < bi)) { bv = sd[w]; bi = si[w]; } }
    idx = bi;
    return bv;
  }
  // "panel" sub-team = the first 4 warps of the CTA, one per scheduler (callers restrict themselves to rank < PW).
  // Its collectives use named barrier 1 and the upper halves of sd/si, never the CTA-wide barrier 0.
  static constexpr int PW = 128, PRU = 2;
  DEVFN void psync() const { asm volatile("bar.sync 1, 128;" ::: "memory"); }
  // v >= 0 or -1 (no candidate): integer max-reductions on the IEEE bit pattern (monotonic for v >= 0)
  DEVFN double pargmax(double v, int &idx) const {
    const unsigned long long key = v < 0.0 ? 0ull : (unsigned long long)__double_as_longlong(v);
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    const bool cand = (hi == mh) && (lo == ml);
    const int widx = (int)__reduce_min_sync(0xffffffffu, cand ? (unsigned)idx : 0x7fffffffu);
    if ((rank & 31) == 0) { sd[32 + (rank >> 5)] = __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml)); si[32 + (rank >> 5)] = widx; }
    psync();
    double bv = sd[32]; int bi = si[32];
#pragma unroll
    for (int w = 1; w < 4; w++) { const double ov = sd[32 + w]; const int oi = si[32 + w]; if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; } }
    idx = bi;
    return bv;
  }
  DEVFN int wfetch(int *ctr) const {
    int g = 0;
    if ((rank & 31) == 0) g = atomicAdd(ctr, 1);
    return __shfl_sync(0xffffffffu, g, 0);
  }
  // owner's v[j0..N_) to everyone in the sub-team
  template <int N_>
  DEVFN void pbcast_n(double (&v)[N_], int j0, int owner) const {
    if (rank == owner) {
#pragma unroll
      for (int j = 0; j < N_; j++) if (j >= j0) sd[40 + j] = v[j];
    }
    psync();
#pragma unroll
    for (int j = 0; j < N_; j++) if (j >= j0) v[j] = sd[40 + j];
  }
};
#endif

// vlidort_Taylor.f90:119-150
DEVFN double taylor_series_1(double eps, double delta, double udel, double sm) {
  const int mterms = TAYLOR_ORDER + 1;
  double d[8];
  d[0] = 1.0;
  for (int m = 1; m <= mterms; m++) d[m] = delta * d[m - 1] / (double)m;
  double mult = d[1], power = 1.0;
  for (int m = 2; m <= mterms; m++) { power *= eps; mult += power * d[m]; }
  return mult * udel * sm;
}
// vlidort_Taylor.f90:152-270
DEVFN double taylor_series_2(double eps, double y, double delta, double fac1, double fac2, double sm) {
  int mterms = TAYLOR_ORDER + 1;
  const bool ysmall = fabs(y) < TAYLOR_SMALL;
  if (ysmall) mterms++;
  double d[8], ac[8], cc[8];
  d[0] = 1.0;
  for (int m = 1; m <= mterms; m++) d[m] = delta * d[m - 1] / (double)m;
  if (ysmall) {
    double power = 1.0, power2 = 1.0, mult = d[2];
    for (int m = 3; m <= mterms; m++) { power = power * (eps - y); power2 = power - y * power2; mult += d[m] * power2; }
    return mult * fac1 * sm;
  }
  const double y1 = 1.0 / y;
  ac[0] = 1.0;
  for (int m = 1; m <= mterms; m++) ac[m] = y1 * ac[m - 1];
  cc[0] = 1.0;
  for (int m = 1; m <= mterms; m++) { cc[m] = 0.0; for (int j = 0; j <= m; j++) cc[m] += ac[j] * d[m - j]; }
  double power = 1.0, mult = fac1 * ac[1] - fac2 * cc[1];
  for (int m = 2; m <= mterms; m++) { power = eps * power; mult += power * (fac1 * ac[m] - fac2 * cc[m]); }
  return mult * sm * y1;
}

// Generalized spherical functions of Fourier order m at cosine x, l = m..NM, already multiplied by
// PINORM(l,m) = sqrt((l-m)!/(l+m)!):  out[3*l+0] = P^l_m (the (1,1)/(4,4) element), out[3*l+1] = R (2,2)=(3,3),
// out[3*l+2] = T (2,3)=(3,2).  Same recurrences and starting values as VLIDORT_PIMATRIX_SETUP_OMP
// (vlidort_miscsetups.f90:2380-2840); the m>2 diagonal start is evaluated as a product instead of being
// carried over from the previous Fourier call.
HDFN void pi_prt(int m, int NM, double x, double *out) {
  const double RC20 = 0.25 * sqrt(6.0), RC21 = 2.0 * RC20;
  const double xsq = x * x, umx = 1.0 - xsq;
  const bool xone = (x == 1.0);
  const double rt = xone ? 0.0 : sqrt(umx);
  double Pm1 = 0, Rm1 = 0, Tm1 = 0, P = 0, R = 0, Tt = 0;     // values at l-1 and l
  double phi;                                                   // PHI(l,m) running product
  {
    double fac = 1.0;
    for (int l = 1; l <= 2 * m; l++) fac *= (double)l;
    phi = 1.0 / fac;
  }
  int l0;
  if (m == 0) {
    // l = 0, 1, 2 explicit
    out[0] = 1.0; out[1] = 0.0; out[2] = 0.0;                   // phi(0,0) = 1
    if (NM >= 1) { out[3] = x; out[4] = 0.0; out[5] = 0.0; }
    if (NM < 2) return;
    Pm1 = x; Rm1 = 0.0; Tm1 = 0.0;
    P = 0.5 * (3.0 * xsq - 1.0); R = RC20 * umx; Tt = 0.0;
    out[6] = P; out[7] = R; out[8] = 0.0;
    for (int l = 2; l <= NM - 1; l++) {
      const double dl = l, x2 = x * (2.0 * dl + 1.0);
      const double Pn = (x2 * P - dl * Pm1) / (dl + 1.0);
      const double Rn = (x2 * R - (l >= 3 ? sqrt(dl * dl - 4.0) : 0.0) * Rm1) / sqrt((dl + 3.0) * (dl - 1.0));
      Pm1 = P; Rm1 = R; P = Pn; R = Rn;
      out[3 * (l + 1) + 0] = P; out[3 * (l + 1) + 1] = R; out[3 * (l + 1) + 2] = 0.0;
    }
    return;
  }
  auto xlm1 = [&](int l) { return 1.0 / (double)(l + 1 - m); };
  auto xlm2 = [&](int l) { return l == 1 ? 0.0 : 1.0 / (sqrt((double)((l + 3) * (l - 1))) * (double)(l + 1 - m) / (double)(l + 1)); };
  auto ylm1 = [&](int l) { return (double)(l + m); };
  auto ylm2 = [&](int l) { return (double)(l + m) * (l >= 3 ? sqrt((double)(l * l - 4)) : 0.0) / (double)l; };
  auto zlm = [&](int l) { return l >= 2 ? (double)m * 2.0 * (2.0 * l + 1.0) / (double)(l + 1) / (double)l : 0.0; };
  if (m == 1) {
    // l = 1
    Pm1 = rt; Rm1 = 0.0; Tm1 = 0.0;
    out[3 * 1 + 0] = Pm1 * sqrt(phi); out[3 * 1 + 1] = 0.0; out[3 * 1 + 2] = 0.0;
    if (NM < 2) return;
    phi *= (double)(2 - m) / (double)(2 + m);
    P = 3.0 * rt * x; R = -RC21 * rt * x; Tt = RC21 * rt;
    { const double pn = sqrt(phi); out[6] = P * pn; out[7] = R * pn; out[8] = Tt * pn; }
    l0 = 2;
  } else {
    double p11 = 3.0 * umx, pkm = xone ? RC21 : RC21 * umx;
    for (int mm = 3; mm <= m; mm++) {
      const double H1 = (double)(2 * mm - 1), H2 = (double)mm * H1 / sqrt((double)((mm + 2) * (mm - 2)));
      p11 = H1 * rt * p11;
      pkm = xone ? 0.0 : H2 * rt * pkm;
    }
    P = p11;
    if (xone) { R = 2.0 * pkm; Tt = -2.0 * pkm; } else { R = pkm * (1.0 + xsq) / umx; Tt = pkm * (-2.0 * x) / umx; }
    { const double pn = sqrt(phi); out[3 * m + 0] = P * pn; out[3 * m + 1] = R * pn; out[3 * m + 2] = Tt * pn; }
    if (m >= NM) return;
    // l = m -> m+1 (no l-1 term)
    const double x2 = x * (2.0 * m + 1.0), z = zlm(m);
    const double Pn = x2 * P * xlm1(m), Rn = (x2 * R + z * Tt) * xlm2(m), Tn = (x2 * Tt + z * R) * xlm2(m);
    Pm1 = P; Rm1 = R; Tm1 = Tt; P = Pn; R = Rn; Tt = Tn;
    phi *= (double)(m + 1 - m) / (double)(m + 1 + m);
    { const double pn = sqrt(phi); out[3 * (m + 1) + 0] = P * pn; out[3 * (m + 1) + 1] = R * pn; out[3 * (m + 1) + 2] = Tt * pn; }
    l0 = m + 1;
  }
  for (int l = l0; l <= NM - 1; l++) {
    const double x2 = x * (2.0 * l + 1.0), z = zlm(l);
    const double Pn = (x2 * P - ylm1(l) * Pm1) * xlm1(l);
    const double Rn = (x2 * R - ylm2(l) * Rm1 + z * Tt) * xlm2(l);
    const double Tn = (x2 * Tt - ylm2(l) * Tm1 + z * R) * xlm2(l);
    Pm1 = P; Rm1 = R; Tm1 = Tt; P = Pn; R = Rn; Tt = Tn;
    phi *= (double)(l + 1 - m) / (double)(l + 1 + m);
    const double pn = sqrt(phi);
    out[3 * (l + 1) + 0] = P * pn; out[3 * (l + 1) + 1] = R * pn; out[3 * (l + 1) + 2] = Tt * pn;
  }
}

}  // namespace vlk
