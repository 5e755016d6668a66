// sdb_kernels.cuh -- device code of libsdb (sm_100a).  One header, compiled two ways:
//   * ahead of time by nvcc into libsdb.so for the named configurations (sdb_aot.cu);
//   * at run time by NVRTC for any other (drift, diffusion, d, m) and for user CUDA source
//     (sdb_capi.cu embeds this file verbatim).
// It therefore includes no host headers.
//
// Work decomposition: ONE THREAD PER SAMPLE PATH.  A path's state, its Wiener increments and
// its repeated integrals live in that thread's registers (local arrays, fully unrolled for
// small d, m); consecutive threads own consecutive paths so every global access -- the
// decimated trajectory stores, host-supplied dW / I loads -- is unit-stride across a warp.
// Operator parameters (A, B_k) are passed BY VALUE as __grid_constant__ kernel parameters
// when they fit in the 32 KB parameter space: fully unrolled loops then address them with
// compile-time offsets and ptxas folds them into DFMA constant-bank operands, so the matrix
// never costs a load instruction.
#pragma once

#ifdef __CUDACC_RTC__
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef long long int64_t;
#else
#include <stdint.h>
#endif

#include "sdb_tables.cuh"

#define SDB_DEV __device__ __forceinline__
#define SDB_HD __host__ __device__ __forceinline__

#ifndef SDB_BLOCK
#define SDB_BLOCK 128
#endif

// ---- ids shared with include/sdb.h --------------------------------------------------------
#define SDB_SCHEME_EULER 0
#define SDB_SCHEME_HEUN 1
#define SDB_SCHEME_SRK2 2
#define SDB_SCHEME_KP2IS 3
#define SDB_SCHEME_R2 4  // Burrage & Burrage (1996) two-stage SRK "R2" (Stratonovich, commutative noise)
#define SDB_NOISE_HOST 0
#define SDB_NOISE_KPW 1
#define SDB_NOISE_WIK 2
#define SDB_NOISE_COMM 3  // host-side only: kernels see KPW with n_terms = 0
#define SDB_FLAG_Y0_BROADCAST 1
#define SDB_FLAG_STRATONOVICH 2

// ---- path functionals ("observers", include/sdb.h sdb_observer) -----------------------------
// Quantities of the FULL-resolution trajectory that a decimated store cannot give back: first
// passage times, running extrema, time integrals -- evaluated at every grid point inside the step
// loop.  Compiled in only when SDB_OBSERVE is 1 (NVRTC variants; the ahead-of-time kernels are
// unchanged).
#ifndef SDB_OBSERVE
#define SDB_OBSERVE 0
#endif
// Philox step counter of interval n.  Resumed segments (sdb_integrate_segment, step0 != 0) offset it;
// like the observers this is compiled in only on request (NVRTC variants), so the ahead-of-time
// kernels do not pay for it (the extra add cost the headline kernel 1.2 %).
#ifndef SDB_RESUME
#define SDB_RESUME 0
#endif
#if SDB_RESUME
#define SDB_STEP(n) ((uint32_t)(n) + a.obs.step0)
#else
#define SDB_STEP(n) ((uint32_t)(n))
#endif
#define SDB_MAX_OBS 4
#define SDB_OBS_FIRST_ABOVE 0
#define SDB_OBS_FIRST_BELOW 1
#define SDB_OBS_MAX 2
#define SDB_OBS_MIN 3
#define SDB_OBS_INTEGRAL 4
#define SDB_OBS_LAGPROD 5    // sum over grid points n >= L of y[comp](t_n) y[comp](t_{n-L}), L = (int)level
#define SDB_OBS_MAXLAG 32    // longest lag (grid steps) of a lagged-product observer: ring in local memory
struct SdbObsArgs {
  int n;
  int kind[SDB_MAX_OBS];
  int comp[SDB_MAX_OBS];
  // Philox step counter of the first interval (resumed segments, SDB_RESUME builds).  It sits in the
  // alignment hole of this block on purpose: the size and layout of the kernel parameters -- and with
  // them the register allocation of the ahead-of-time kernels -- stay what they were.
  uint32_t step0;
  double level[SDB_MAX_OBS];
  double* out;  // [observer][path] through strides
  int64_t sO, sP;
};

// ---- kernel argument block (plain data; mirrored by the host in sdb_capi.cu) ---------------
struct SdbLoopArgs {
  int64_t P;
  int64_t path_id0;
  uint64_t seed;
  int N;
  int save_every;
  int n_terms;
  int flags;
  double h_global;  // (t[N-1]-t[0])/(N-1): integrate.py:228,285,480,594
  double rtol;      // KP2iS implicit-solve tolerance
  const double* tspan;  // device [N]
  const double* y0;
  int64_t y0_sI, y0_sP;
  double* y;
  int64_t y_sS, y_sI, y_sP;
  const double* dW;
  int64_t dW_sN, dW_sJ, dW_sP;
  const double* IJ;
  int64_t IJ_sN, IJ_sR, IJ_sC, IJ_sP;
  const double* kp;  // device [3*D]: gam, al1, al2
  int64_t* status;   // [4]
  SdbObsArgs obs;    // per-path functionals of the trajectory (SDB_OBSERVE builds) and the step offset
};

struct SdbRepintArgs {
  int method, stratonovich, n_terms, m;
  int64_t steps, paths, path_id0;
  uint64_t seed;
  double h;
  const double* dW; int64_t dW_sN, dW_sJ, dW_sP;
  const double* X; const double* Y; int64_t XY_sT, XY_sN, XY_sJ, XY_sP;
  const double* Gt; int64_t Gt_sN, Gt_sR, Gt_sP;
  double* A; int64_t A_sN, A_sE, A_sP;
  double* IJ; int64_t IJ_sN, IJ_sR, IJ_sC, IJ_sP;
};

// ---- parameter stores --------------------------------------------------------------------
template <int NP>
struct PInline {  // by value in the kernel parameter space (constant bank)
  double v[NP > 0 ? NP : 1];
  SDB_DEV double operator[](int i) const { return v[i]; }
  SDB_DEV const double* ptr() const { return v; }
};
struct PGlobal {  // too large for the parameter space: read through the L1/L2 hierarchy
  const double* v;
  SDB_DEV double operator[](int i) const { return __ldg(v + i); }
  SDB_DEV const double* ptr() const { return v; }
};

// ---- Philox4x32-10 (Salmon et al. 2011) and the normal transform -------------------------
#define SDB_PHILOX_M0 0xD2511F53u
#define SDB_PHILOX_M1 0xCD9E8D57u
#define SDB_PHILOX_W0 0x9E3779B9u
#define SDB_PHILOX_W1 0xBB67AE85u

// Round keys k + r W (warp-uniform): built once per kernel so they live in uniform registers and a
// round is exactly 2 IMAD.WIDE + 2 LOP3.
struct SdbPhiloxKey {
  uint32_t a[10], b[10];
  SDB_HD explicit SdbPhiloxKey(uint64_t seed) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      a[r] = (uint32_t)seed + (uint32_t)r * SDB_PHILOX_W0;
      b[r] = (uint32_t)(seed >> 32) + (uint32_t)r * SDB_PHILOX_W1;
    }
  }
};

SDB_HD void sdb_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                              const SdbPhiloxKey& key, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)SDB_PHILOX_M0 * c0;
    const uint64_t p1 = (uint64_t)SDB_PHILOX_M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ key.a[r];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.b[r];
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ---- keyed expansion of the per-(path, step) Philox block ("MX2") -----------------------------
// IMAD.WIDE, the 64-bit product of Philox, occupies the FP64 pipe's issue resource on sm_100a
// (profiles/r01_microbench.md), so only TWO Philox4x32-10 blocks are computed per (path, step):
//   slot 0                 -> the block itself (counter word 3 = 0), as before;
//   slot s >= 1            -> MX2(K, s) with K = Philox4x32-10 at counter word 3 = SDB_RNG_KEYSLOT.
// MX2 is two rounds of 32-bit multiply-xorshift over four coupled words (12 IMAD + 16 ALU
// instructions, none on the FP64 pipe); K is never output.  It ends on a multiply-add, whose HIGH
// bits are the well-mixed ones: the normal transform below takes the leading mantissa bits and the
// octant bits from the top of w0 / w2 and uses w1 / w3 whole as the low mantissa words (their weak
// low bits are the last bits of a 52-bit uniform).  Spec restated in oracle/philox_oracle.py;
// avalanche / correlation / normal-distribution tests in tests/test_philox.py.
#define SDB_RNG_KEYSLOT 0xFFFFFFFFu
#define SDB_MX_G0 0x9E3779B1u
#define SDB_MX_G1 0x85EBCA77u
#define SDB_MX_G2 0xC2B2AE3Du
#define SDB_MX_G3 0x27D4EB2Fu
#define SDB_MX_C0 0xED5AD4BBu
#define SDB_MX_C1 0xAC4C1B51u
#define SDB_MX_C2 0x31848BABu
#define SDB_MX_C3 0x7FEB352Du
SDB_HD void sdb_mx2_block(const uint32_t (&K)[4], uint32_t s, uint32_t (&o)[4]) {
  uint32_t x0 = K[0] + s * SDB_MX_G0, x1 = K[1] + s * SDB_MX_G1, x2 = K[2] + s * SDB_MX_G2,
           x3 = K[3] + s * SDB_MX_G3;
  x0 ^= x0 >> 15; x1 ^= x1 >> 15; x2 ^= x2 >> 15; x3 ^= x3 >> 15;
  uint32_t y0 = x0 * SDB_MX_C0 + x1, y1 = x1 * SDB_MX_C1 + x2, y2 = x2 * SDB_MX_C2 + x3,
           y3 = x3 * SDB_MX_C3 + x0;
  y0 ^= y0 >> 13; y1 ^= y1 >> 13; y2 ^= y2 >> 13; y3 ^= y3 >> 13;
  o[0] = y0 * SDB_MX_C1 + y1; o[1] = y1 * SDB_MX_C2 + y2; o[2] = y2 * SDB_MX_C3 + y3; o[3] = y3 * SDB_MX_C0 + y0;
}

// Counter layout: (path low, path high, step, slot); key = seed.  One 128-bit block yields one
// Box-Muller pair (spec restated in oracle/philox_oracle.py):
//   a = (w0 >> 12) 2^32 + w1 (52 bits: the top 20 of w0, all of w1), u1 = (2 a + 1) 2^-53 in (0,1)
//   rad = sqrt(-2 ln u1);  t = (((w2 >> 9) & (2^20-1)) 2^32 + w3) 2^-52 in [0,1);  (c, s) = cos, sin(pi t / 4)
//   octant bits o = w2 >> 29: swap = o&1, neg0 = o&2, neg1 = o&4
//   z0 = +-rad * (swap ? s : c),  z1 = +-rad * (swap ? c : s)        -- a uniform angle, by symmetry
// Normal number q of a path-step is element (q & 1) of slot (q >> 1):
//   dW_j -> q = j;  X_k[j] -> m + 2(k-1)m + j;  Y_k[j] -> m + (2(k-1)+1)m + j;  Gtail_r -> m + 2nm + r.
//
// The transform is written for the FP64 pipe budget (32 DFMA-class instructions per pair, the
// rest on the integer / XU pipes): ln u1 by a 128-entry table (1/c, -2 ln c) + degree-6 log1p,
// sqrt by MUFU.RSQ64H + one Newton step + one residual correction, sin/cos by degree-6
// polynomials in t^2.  Accuracy ~1e-16 relative (tests compare with numpy's libm).
#define SDB_ONE_M53 0.99999999999999988898  // 1 - 2^-53 (0x3FEFFFFFFFFFFFFF)
SDB_DEV double sdb_rsqrt_approx(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}

SDB_DEV void sdb_normal_pair(const uint32_t (&w)[4], const double2* __restrict__ tab, double& z0,
                             double& z1) {
  // ---- radius -----------------------------------------------------------------------------------
  // u1 = (2 a + 1) 2^-53 with a = (w0 >> 12):w1: 1 + a 2^-52 is a double with those mantissa words, and
  // subtracting 1 - 2^-53 is exact, so one DADD converts; the rest works on the high word only.
  const double x = __hiloint2double((int)(0x3FF00000u | (w[0] >> 12)), (int)w[1]) - SDB_ONE_M53;
  const int xh = __double2hiint(x);
  const int tmp = xh - 0x3FE60000;  // high word of SDB_LOG_OFF
  const int i = (tmp >> 13) & 127;
  const int k = tmp >> 20;          // u1 = 2^k z,  z in [0.6875, 1.375)
  const double z = __hiloint2double(xh - (tmp & (int)0xFFF00000), __double2loint(x));
  const double2 tc = tab[i];  // {1/c, -2 ln c}
  const double r = fma(z, tc.x, -1.0);
  double q = fma(r, 1.0 / 3.0, -0.4);
  q = fma(r, q, 0.5);
  q = fma(r, q, -2.0 / 3.0);
  q = fma(r, q, 1.0);
  q = fma(r, q, -2.0);
  double L = fma((double)k, SDB_M2LN2, tc.y);
  L = fma(r, q, L);  // -2 ln u1 = -2 k ln 2 - 2 ln c - 2 r + r^2 - (2/3) r^3 + ... > 0
  // sqrt(L) = g / sqrt(1 - e), g = L y, e = 1 - L y^2 (|e| < 2^-19 after MUFU.RSQ64H): three terms of
  // the binomial series leave 5 e^3 / 16 < 2^-58
  const double y = sdb_rsqrt_approx(L);
  const double g = L * y;
  const double e = fma(-g, y, 1.0);
  const double rad = fma(g * e, fma(e, 0.375, 0.5), g);
  // ---- angle ------------------------------------------------------------------------------------
  const double t = __hiloint2double((int)(0x3FF00000u | ((w[2] >> 9) & 0x000FFFFFu)), (int)w[3]) - 1.0;
  const double t2 = t * t;
  double sn = fma(t2, SDB_SIN6, SDB_SIN5);
  sn = fma(t2, sn, SDB_SIN4);
  sn = fma(t2, sn, SDB_SIN3);
  sn = fma(t2, sn, SDB_SIN2);
  sn = fma(t2, sn, SDB_SIN1);
  sn = fma(t2, sn, SDB_SIN0);
  sn *= t;
  double cs = fma(t2, SDB_COS6, SDB_COS5);
  cs = fma(t2, cs, SDB_COS4);
  cs = fma(t2, cs, SDB_COS3);
  cs = fma(t2, cs, SDB_COS2);
  cs = fma(t2, cs, SDB_COS1);
  cs = fma(t2, cs, SDB_COS0);
  const bool swap = (w[2] >> 29) & 1u;
  const double a0 = swap ? sn : cs;
  const double b0 = swap ? cs : sn;
  const uint32_t s0 = (w[2] << 1) & 0x80000000u;  // bit 30 -> sign
  const uint32_t s1 = w[2] & 0x80000000u;         // bit 31 -> sign
  const double p0 = rad * a0, p1 = rad * b0;
  z0 = __hiloint2double(__double2hiint(p0) ^ (int)s0, __double2loint(p0));
  z1 = __hiloint2double(__double2hiint(p1) ^ (int)s1, __double2loint(p1));
}

// The 128-bit random block of (seed, path, step, slot): host and device (sdb_random_block in
// include/sdb.h exposes it to the tests).
SDB_HD void sdb_random_block(const SdbPhiloxKey& key, uint64_t path, uint32_t step, uint32_t slot,
                             uint32_t (&out)[4]) {
  if (slot == 0u) {
    sdb_philox4x32_10((uint32_t)path, (uint32_t)(path >> 32), step, 0u, key, out);
  } else {
    uint32_t K[4];
    sdb_philox4x32_10((uint32_t)path, (uint32_t)(path >> 32), step, SDB_RNG_KEYSLOT, key, K);
    sdb_mx2_block(K, slot, out);
  }
}

struct SdbStream {
  const SdbPhiloxKey& key;
  uint32_t p_lo, p_hi, step;
  const double2* tab;  // log table: global (sdb_logtab) or a shared-memory copy
  uint32_t K[4];       // expansion key of this (path, step); only built when slots >= 1 are read
  SDB_DEV SdbStream(const SdbPhiloxKey& key_, uint64_t path, uint32_t step_,
                    const double2* tab_ = reinterpret_cast<const double2*>(sdb_logtab),
                    bool expand = true)
      : key(key_), p_lo((uint32_t)path), p_hi((uint32_t)(path >> 32)), step(step_), tab(tab_) {
    if (expand) {
      sdb_philox4x32_10(p_lo, p_hi, step, SDB_RNG_KEYSLOT, key, K);
    } else {
      K[0] = K[1] = K[2] = K[3] = 0u;
    }
  }
  SDB_DEV void block(uint32_t slot, uint32_t (&r)[4]) const {
    if (slot == 0u) {
      sdb_philox4x32_10(p_lo, p_hi, step, 0u, key, r);
    } else {
      sdb_mx2_block(K, slot, r);
    }
  }
  SDB_DEV void pair(uint32_t slot, double& z0, double& z1) const {
    uint32_t r[4];
    block(slot, r);
    sdb_normal_pair(r, tab, z0, z1);
  }
};

// Sequential reader of normals q0, q0+1, ... (keeps the odd element of a pair for the next call).
struct SdbNormalSeq {
  const SdbStream& st;
  uint32_t q;
  double spare;
  SDB_DEV SdbNormalSeq(const SdbStream& s, uint32_t q0) : st(s), q(q0), spare(0.0) {
    if (q & 1u) { double z0; st.pair(q >> 1, z0, spare); }
  }
  SDB_DEV double next() {
    double z;
    if (q & 1u) {
      z = spare;
    } else {
      st.pair(q >> 1, z, spare);
    }
    ++q;
    return z;
  }
};

// ---- repeated integrals, elementwise (oracle/sde_oracle.py: ikpw_from_normals,
// iwik_from_normals; reference sdeint/wiener.py:67-106 and :206-282) ---------------------------
#define SDB_INV_2PI 0.15915494309189535

SDB_DEV double sdb_a_tail(int n) {  // wiener.py:229-231
  double s = 0.0;
  for (int k = 1; k <= n; ++k) s += 1.0 / ((double)k * (double)k);
  return 1.6449340668482264 - s;  // pi^2/6
}

// Accumulate the Levy-area series term k into the strict upper triangle At (K_m row order).
template <int M>
SDB_DEV void sdb_levy_accumulate(const double (&X)[M], const double (&Z)[M], double inv_k,
                                 double (&At)[M * (M - 1) / 2 + 1]) {
  int r = 0;
#pragma unroll
  for (int i = 0; i < M; ++i) {
#pragma unroll
    for (int j = i + 1; j < M; ++j) {
      At[r] = fma(X[i] * Z[j] - Z[i] * X[j], inv_k, At[r]);
      ++r;
    }
  }
}

// Source of the standard normals: Philox stream or caller-supplied arrays.
template <int M>
struct SdbNormalsPhilox {
  const SdbStream& st;
  int n_terms;
  SDB_DEV void xk(int k, double (&X)[M]) const {  // k = 1..n
    SdbNormalSeq s(st, (uint32_t)(M + 2 * (k - 1) * M));
#pragma unroll
    for (int j = 0; j < M; ++j) X[j] = s.next();
  }
  SDB_DEV void yk(int k, double (&Y)[M]) const {
    SdbNormalSeq s(st, (uint32_t)(M + (2 * (k - 1) + 1) * M));
#pragma unroll
    for (int j = 0; j < M; ++j) Y[j] = s.next();
  }
  SDB_DEV void tail(double (&G)[M * (M - 1) / 2 + 1]) const {
    SdbNormalSeq s(st, (uint32_t)(M + 2 * n_terms * M));
#pragma unroll 1
    for (int r = 0; r < M * (M - 1) / 2; ++r) G[r] = s.next();
  }
};

template <int M>
struct SdbNormalsHost {
  const double* X; const double* Y; int64_t sT, sJ;  // already offset to (step, path)
  const double* Gt; int64_t sR;
  SDB_DEV void xk(int k, double (&Xo)[M]) const {
#pragma unroll
    for (int j = 0; j < M; ++j) Xo[j] = X[(k - 1) * sT + j * sJ];
  }
  SDB_DEV void yk(int k, double (&Yo)[M]) const {
#pragma unroll
    for (int j = 0; j < M; ++j) Yo[j] = Y[(k - 1) * sT + j * sJ];
  }
  SDB_DEV void tail(double (&G)[M * (M - 1) / 2 + 1]) const {
#pragma unroll 1
    for (int r = 0; r < M * (M - 1) / 2; ++r) G[r] = Gt[r * sR];
  }
};

// Kloeden-Platen-Wright series: At = (h/2pi) sum_k (X_k[i] Z_k[j] - Z_k[i] X_k[j]) / k, i<j.
template <int M, class NS>
SDB_DEV void sdb_levy_series(const double (&dW)[M], double h, int n_terms, const NS& ns,
                             double (&At)[M * (M - 1) / 2 + 1]) {
  constexpr int MM = M * (M - 1) / 2;
#pragma unroll
  for (int r = 0; r < MM + 1; ++r) At[r] = 0.0;
  const double c = sqrt(2.0 / h);
#pragma unroll 1
  for (int k = 1; k <= n_terms; ++k) {
    double X[M], Z[M];
    ns.xk(k, X);
    ns.yk(k, Z);
#pragma unroll
    for (int j = 0; j < M; ++j) Z[j] = fma(c, dW[j], Z[j]);
    sdb_levy_accumulate<M>(X, Z, 1.0 / (double)k, At);
  }
  const double s = h * SDB_INV_2PI;
#pragma unroll
  for (int r = 0; r < MM; ++r) At[r] *= s;
}

// Wiktorsson tail correction applied in O(m^2): At += (h/2pi) sqrt(a_n) sqrtSigma g.
template <int M, class NS>
SDB_DEV void sdb_wik_tail(const double (&dW)[M], double h, int n_terms, const NS& ns,
                          double (&At)[M * (M - 1) / 2 + 1]) {
  constexpr int MM = M * (M - 1) / 2;
  double G[MM + 1];
  ns.tail(G);
  double nrm = 0.0;
#pragma unroll
  for (int j = 0; j < M; ++j) nrm = fma(dW[j], dW[j], nrm);
  const double rad = sqrt(1.0 + nrm / h);
  double u[M];
#pragma unroll
  for (int j = 0; j < M; ++j) u[j] = 0.0;
  {
    int r = 0;
#pragma unroll
    for (int i = 0; i < M; ++i) {
#pragma unroll
      for (int j = i + 1; j < M; ++j) {
        u[i] = fma(G[r], dW[j], u[i]);
        u[j] = fma(-G[r], dW[i], u[j]);
        ++r;
      }
    }
  }
  const double scale = h * SDB_INV_2PI * sqrt(sdb_a_tail(n_terms));
  const double den = 1.0 / (1.4142135623730951 * (1.0 + rad));
  const double two_h = 2.0 / h;
  {
    int r = 0;
#pragma unroll
    for (int i = 0; i < M; ++i) {
#pragma unroll
      for (int j = i + 1; j < M; ++j) {
        const double sig = fma(two_h, dW[j] * u[i] - dW[i] * u[j], 2.0 * G[r]);
        const double root = fma(2.0 * rad, G[r], sig) * den;
        At[r] = fma(scale, root, At[r]);
        ++r;
      }
    }
  }
}

// Assemble I (or J) from dW and the area vector.  KPW places +A above the diagonal
// (I[i][j] = dWi dWj/2 + A_ij, wiener.py:106); Wiktorsson's column-major unvec places it BELOW
// (I[j][i] = dWi dWj/2 + Atilde_ij, wiener.py:278-280).
template <int M>
SDB_DEV void sdb_assemble_IJ(const double (&dW)[M], double h, const double (&At)[M * (M - 1) / 2 + 1],
                             bool wik, bool strat, double (&IJ)[M][M]) {
  int r = 0;
#pragma unroll
  for (int i = 0; i < M; ++i) {
    IJ[i][i] = strat ? 0.5 * dW[i] * dW[i] : 0.5 * (dW[i] * dW[i] - h);
#pragma unroll
    for (int j = i + 1; j < M; ++j) {
      const double sym = 0.5 * dW[i] * dW[j];
      const double a = wik ? -At[r] : At[r];
      IJ[i][j] = sym + a;
      IJ[j][i] = sym - a;
      ++r;
    }
  }
}

// Levy areas only (series + optional Wiktorsson tail), K_m row order; At[0] = 0 for M == 1.
template <int M, class NS>
SDB_DEV void sdb_levy_areas(const double (&dW)[M], double h, int n_terms, int method, const NS& ns,
                            double (&At)[M * (M - 1) / 2 + 1]) {
  if (M == 1) {
    At[0] = 0.0;
    return;
  }
  sdb_levy_series<M>(dW, h, n_terms, ns, At);
  if (method == SDB_NOISE_WIK) sdb_wik_tail<M>(dW, h, n_terms, ns, At);
}

template <int M, class NS>
SDB_DEV void sdb_repeated_integrals(const double (&dW)[M], double h, int n_terms, int method,
                                    bool strat, const NS& ns, double (&At)[M * (M - 1) / 2 + 1],
                                    double (&IJ)[M][M]) {
  sdb_levy_areas<M>(dW, h, n_terms, method, ns, At);
  if (M == 1) {
    if (method == SDB_NOISE_KPW) {  // Ikpw still draws its 2n normals; A == 0
      IJ[0][0] = strat ? 0.5 * dW[0] * dW[0] : 0.5 * (dW[0] * dW[0] - h);
    } else {                         // Iwik shortcut, wiener.py:259-260
      IJ[0][0] = strat ? (dW[0] * dW[0] - h) / 2.0 + 0.5 * h : (dW[0] * dW[0] - h) / 2.0;
    }
    return;
  }
  sdb_assemble_IJ<M>(dW, h, At, method == SDB_NOISE_WIK, strat, IJ);
}

// ---- drift / diffusion operators -------------------------------------------------------------
// Interface:  Drift::eval(params, y[D], t, out[D]);  Diff::col(params, k, y[D], t, out[D]).
template <int D_, class PS>
struct DriftLinear {
  static constexpr int D = D_;
  static constexpr bool LINEAR = true;  // the Jacobian is A: KP2iS needs no finite differences
  typedef PS Params;
  static SDB_DEV double jac(const PS& p, int i, int j) { return p[i * D + j]; }
  static SDB_DEV void eval(const PS& p, const double (&y)[D], double, double (&out)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double acc = p[i * D] * y[0];
#pragma unroll
      for (int j = 1; j < D; ++j) acc = fma(p[i * D + j], y[j], acc);
      out[i] = acc;
    }
  }
};

template <class PS>
struct DriftKP4446 {  // -(a + c y)(1 - y^2), sdeint/tests/test_integrate.py:93-99
  static constexpr int D = 1;
  static constexpr bool LINEAR = false;
  typedef PS Params;
  static SDB_DEV void eval(const PS& p, const double (&y)[1], double, double (&out)[1]) {
    out[0] = -(p[0] + y[0] * p[1]) * (1.0 - y[0] * y[0]);
  }
};

template <class PS>
struct DriftKPS445 {  // test_integrate.py:263-269
  static constexpr int D = 1;
  static constexpr bool LINEAR = false;
  typedef PS Params;
  static SDB_DEV void eval(const PS& p, const double (&y)[1], double, double (&out)[1]) {
    double s, c;
    sincos(y[0], &s, &c);
    out[0] = -sin(2.0 * y[0]) - 0.25 * sin(4.0 * y[0]) + p[0] * 2.0 * s * c * c * c;
  }
};

template <int D_, int M_, class PS>
struct DiffConst {  // additive noise: G does not depend on y
  static constexpr int D = D_, M = M_;
  static constexpr bool ADDITIVE = true, DIAGONAL = false;
  typedef PS Params;
  static SDB_DEV void col(const PS& p, int k, const double (&)[D], double, double (&out)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) out[i] = p[i * M + k];
  }
};

template <int D_, int M_, class PS>
struct DiffLinearCols {  // g_k = B_k y
  static constexpr int D = D_, M = M_;
  static constexpr bool ADDITIVE = false, DIAGONAL = false;
  typedef PS Params;
  static SDB_DEV void col(const PS& p, int k, const double (&y)[D], double, double (&out)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double acc = p[(k * D + i) * D] * y[0];
#pragma unroll
      for (int j = 1; j < D; ++j) acc = fma(p[(k * D + i) * D + j], y[j], acc);
      out[i] = acc;
    }
  }
};

template <int D_, int M_, class PS>
struct DiffAffineCols {  // g_k = B_k y + c_k: params B[m][d][d] then C[d][m]
  static constexpr int D = D_, M = M_;
  static constexpr bool ADDITIVE = false, DIAGONAL = false;
  typedef PS Params;
  static SDB_DEV void col(const PS& p, int k, const double (&y)[D], double, double (&out)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double acc = p[M * D * D + i * M + k];
#pragma unroll
      for (int j = 0; j < D; ++j) acc = fma(p[(k * D + i) * D + j], y[j], acc);
      out[i] = acc;
    }
  }
};

// g_k linear in y (no offset): differences g_k(a) - g_k(b) = g_k(a - b) exactly
template <class G_> struct SdbGLinear { static constexpr bool value = false; };
template <int D_, int M_, class PS>
struct SdbGLinear<DiffLinearCols<D_, M_, PS>> { static constexpr bool value = true; };

template <int D_, class PS>
struct DiffDiagLinear {  // g_k = b_k y_k e_k: diagonal noise (m = d)
  static constexpr int D = D_, M = D_;
  static constexpr bool ADDITIVE = false, DIAGONAL = true;
  typedef PS Params;
  static SDB_DEV void col(const PS& p, int k, const double (&y)[D], double, double (&out)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) out[i] = i == k ? p[i] * y[i] : 0.0;
  }
};

template <class PS>
struct DiffKP4446 {
  static constexpr int D = 1, M = 1;
  static constexpr bool ADDITIVE = false, DIAGONAL = false;
  typedef PS Params;
  static SDB_DEV void col(const PS& p, int, const double (&y)[1], double, double (&out)[1]) {
    out[0] = p[0] * (1.0 - y[0] * y[0]);
  }
};

template <class PS>
struct DiffKPS445 {
  static constexpr int D = 1, M = 1;
  static constexpr bool ADDITIVE = false, DIAGONAL = false;
  typedef PS Params;
  static SDB_DEV void col(const PS& p, int, const double (&y)[1], double, double (&out)[1]) {
    const double c = cos(y[0]);
    out[0] = p[0] * c * c;
  }
};

#ifdef SDB_USER_DRIFT
__device__ void sde_drift(const double* y, double t, const double* p, double* out);
template <int D_, class PS>
struct DriftUser {
  static constexpr int D = D_;
  static constexpr bool LINEAR = false;
  typedef PS Params;
  static SDB_DEV void eval(const PS& p, const double (&y)[D], double t, double (&out)[D]) {
    sde_drift(y, t, p.ptr(), out);
  }
};
#endif
#ifdef SDB_USER_DIFF
__device__ void sde_diffusion_col(int k, const double* y, double t, const double* p, double* out);
template <int D_, int M_, class PS, bool DIAG = false>
struct DiffUser {
  static constexpr int D = D_, M = M_;
  static constexpr bool ADDITIVE = false, DIAGONAL = DIAG;
  typedef PS Params;
  static SDB_DEV void col(const PS& p, int k, const double (&y)[D], double t, double (&out)[D]) {
    sde_diffusion_col(k, y, t, p.ptr(), out);
  }
};
#endif

// ---- per-step noise for the loop kernels ---------------------------------------------------------
// Diagonal noise: only I_kk = (dW_k^2 - h)/2 (J_kk = dW_k^2/2) enters the schemes; no Levy areas.
template <int M>
SDB_DEV void sdb_diag_IJ(const SdbLoopArgs& a, const double (&dW)[M], double (&IJ)[M][M]) {
  const bool strat = (a.flags & SDB_FLAG_STRATONOVICH) != 0;
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int k = 0; k < M; ++k)
      IJ[j][k] = j != k ? 0.0 : (strat ? 0.5 * dW[j] * dW[j] : 0.5 * (dW[j] * dW[j] - a.h_global));
}

template <int M, int NOISE, bool NEED_IJ, bool DIAG>
SDB_DEV void sdb_step_noise(const SdbLoopArgs& a, int64_t p, int n, double (&dW)[M],
                            double (&IJ)[M][M]) {
  if (NOISE == SDB_NOISE_HOST) {
    const double* w = a.dW + (int64_t)n * a.dW_sN + p * a.dW_sP;
#pragma unroll
    for (int j = 0; j < M; ++j) dW[j] = w[j * a.dW_sJ];
    if (NEED_IJ && DIAG && a.IJ == nullptr) {
      sdb_diag_IJ<M>(a, dW, IJ);
    } else if (NEED_IJ) {
      const double* q = a.IJ + (int64_t)n * a.IJ_sN + p * a.IJ_sP;
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int k = 0; k < M; ++k) IJ[j][k] = q[j * a.IJ_sR + k * a.IJ_sC];
    }
  } else {
    const SdbPhiloxKey key(a.seed);
    // slots >= 1 (the keyed expansion) are read for m > 2 and for the Levy-series normals
    const SdbStream st(key, (uint64_t)(a.path_id0 + p), SDB_STEP(n),
                       reinterpret_cast<const double2*>(sdb_logtab), M > 2 || (NEED_IJ && !DIAG));
    const double rh = sqrt(a.h_global);  // deltaW(N-1, m, h_global): integrate.py:483
    SdbNormalSeq s(st, 0u);
#pragma unroll
    for (int j = 0; j < M; ++j) dW[j] = rh * s.next();
    if (NEED_IJ && DIAG) {
      sdb_diag_IJ<M>(a, dW, IJ);
    } else if (NEED_IJ) {
      double At[M * (M - 1) / 2 + 1];
      const SdbNormalsPhilox<M> ns{st, a.n_terms};
      sdb_repeated_integrals<M>(dW, a.h_global, a.n_terms, NOISE,
                                (a.flags & SDB_FLAG_STRATONOVICH) != 0, ns, At, IJ);
    }
  }
}

// On-device noise for the large-system SRK2 path: dW and the Levy-area vector only -- I / J is never
// materialised (m^2 doubles of local memory); the step loop applies it to G_n row by row from dW and
// the areas, which stay in registers.
template <int M, int NOISE>
SDB_DEV void sdb_step_noise_areas(const SdbLoopArgs& a, int64_t p, int n, double (&dW)[M],
                                  double (&At)[M * (M - 1) / 2 + 1]) {
  const SdbPhiloxKey key(a.seed);
  const SdbStream st(key, (uint64_t)(a.path_id0 + p), SDB_STEP(n));
  const double rh = sqrt(a.h_global);
  SdbNormalSeq s(st, 0u);
#pragma unroll
  for (int j = 0; j < M; ++j) dW[j] = rh * s.next();
  const SdbNormalsPhilox<M> ns{st, a.n_terms};
  sdb_levy_areas<M>(dW, a.h_global, a.n_terms, NOISE, ns, At);
}

template <int D>
SDB_DEV void sdb_store(const SdbLoopArgs& a, int64_t p, int64_t s, const double (&Y)[D]) {
  double* o = a.y + s * a.y_sS + p * a.y_sP;
#pragma unroll
  for (int i = 0; i < D; ++i) o[i * a.y_sI] = Y[i];
}

template <int D>
SDB_DEV void sdb_finish(const SdbLoopArgs& a, int64_t p, const double (&Y)[D]) {
  bool ok = true;
#pragma unroll
  for (int i = 0; i < D; ++i) ok = ok && isfinite(Y[i]);
  if (!ok && a.status) {
    atomicAdd((unsigned long long*)&a.status[0], 1ull);
    atomicMin((long long*)&a.status[1], (long long)(a.path_id0 + p));
  }
}

// Observer state of one path: registers only, component picked by a select chain (a runtime index
// into Y would move the whole state to local memory).
template <int D>
struct SdbObs {
  double v[SDB_MAX_OBS], prev[SDB_MAX_OBS];
  // lagged products (autocovariance at FULL time resolution, the spectral check the reference leaves
  // as a to-do, test_integrate.py:455-456): the last L values of the component, written round-robin
  double ring[SDB_MAX_OBS][SDB_OBS_MAXLAG];
  int pos[SDB_MAX_OBS], seen[SDB_MAX_OBS];
  static SDB_DEV double pick(const double (&Y)[D], int c) {
    double r = Y[0];
#pragma unroll
    for (int i = 1; i < D; ++i) r = (i == c) ? Y[i] : r;
    return r;
  }
  SDB_DEV void init(const SdbLoopArgs& a, const double (&Y)[D]) {
    const double t0 = a.tspan[0];
#pragma unroll
    for (int o = 0; o < SDB_MAX_OBS; ++o) {
      v[o] = 0.0;
      prev[o] = 0.0;
      if (o < a.obs.n) {
        const double y = pick(Y, a.obs.comp[o]);
        const int kind = a.obs.kind[o];
        const double inf = __longlong_as_double(0x7ff0000000000000ll);
        prev[o] = y;
        if (kind == SDB_OBS_FIRST_ABOVE) v[o] = y >= a.obs.level[o] ? t0 : inf;
        else if (kind == SDB_OBS_FIRST_BELOW) v[o] = y <= a.obs.level[o] ? t0 : inf;
        else if (kind == SDB_OBS_INTEGRAL) v[o] = 0.0;
        else if (kind == SDB_OBS_LAGPROD) {
          const int L = (int)a.obs.level[o];
          v[o] = L == 0 ? y * y : 0.0;
          ring[o][0] = y;
          pos[o] = L > 1 ? 1 : 0;
          seen[o] = 1;
        } else v[o] = y;
      }
    }
  }
  // after step n: Y = y(t_{n+1})
  SDB_DEV void step(const SdbLoopArgs& a, int n, const double (&Y)[D]) {
    if (a.obs.n == 0) return;
    const double t0 = a.tspan[n], t1 = a.tspan[n + 1];
#pragma unroll
    for (int o = 0; o < SDB_MAX_OBS; ++o) {
      if (o < a.obs.n) {
        const double y = pick(Y, a.obs.comp[o]);
        const int kind = a.obs.kind[o];
        if (kind == SDB_OBS_FIRST_ABOVE) v[o] = (y >= a.obs.level[o] && t1 < v[o]) ? t1 : v[o];
        else if (kind == SDB_OBS_FIRST_BELOW) v[o] = (y <= a.obs.level[o] && t1 < v[o]) ? t1 : v[o];
        else if (kind == SDB_OBS_MAX) v[o] = y > v[o] ? y : v[o];
        else if (kind == SDB_OBS_MIN) v[o] = y < v[o] ? y : v[o];
        else if (kind == SDB_OBS_LAGPROD) {
          const int L = (int)a.obs.level[o];
          if (L == 0) {
            v[o] = fma(y, y, v[o]);
          } else {
            const int i = pos[o];
            if (seen[o] >= L) v[o] = fma(y, ring[o][i], v[o]);  // ring[i] holds y(t_{n+1-L})
            ring[o][i] = y;
            pos[o] = i + 1 == L ? 0 : i + 1;
            seen[o] += 1;
          }
        }
        else v[o] = fma(0.5 * (t1 - t0), prev[o] + y, v[o]);  // trapezoid
        prev[o] = y;
      }
    }
  }
  SDB_DEV void finish(const SdbLoopArgs& a, int64_t p) const {
#pragma unroll
    for (int o = 0; o < SDB_MAX_OBS; ++o)
      if (o < a.obs.n) a.obs.out[o * a.obs.sO + p * a.obs.sP] = v[o];
  }
};

// Small dense solve J x = b with partial pivoting (KP2iS Newton step); J is destroyed.
template <int D>
SDB_DEV bool sdb_solve(double (&J)[D][D], double (&b)[D]) {
#pragma unroll 1
  for (int c = 0; c < D; ++c) {
    int piv = c;
    double best = fabs(J[c][c]);
    for (int r = c + 1; r < D; ++r) {
      const double v = fabs(J[r][c]);
      if (v > best) { best = v; piv = r; }
    }
    if (best == 0.0) return false;
    if (piv != c) {
      for (int k = 0; k < D; ++k) { const double t = J[c][k]; J[c][k] = J[piv][k]; J[piv][k] = t; }
      const double t = b[c]; b[c] = b[piv]; b[piv] = t;
    }
    const double inv = 1.0 / J[c][c];
    for (int r = c + 1; r < D; ++r) {
      const double f = J[r][c] * inv;
      for (int k = c; k < D; ++k) J[r][k] = fma(-f, J[c][k], J[r][k]);
      b[r] = fma(-f, b[c], b[r]);
    }
  }
  for (int r = D - 1; r >= 0; --r) {
    double acc = b[r];
    for (int k = r + 1; k < D; ++k) acc = fma(-J[r][k], b[k], acc);
    b[r] = acc / J[r][r];
  }
  return true;
}

// ---- the step loops ---------------------------------------------------------------------------------
// SCHEME_EULER : integrate.py:235-239      SCHEME_HEUN : integrate.py:292-302
// SCHEME_SRK2  : integrate.py:494-522      SCHEME_KP2IS: integrate.py:613-645
template <int SCHEME, class F, class G, int NOISE>
SDB_DEV void sdb_loop_body(const SdbLoopArgs& a, const typename F::Params& fp,
                           const typename G::Params& gp) {
  constexpr int D = G::D;
  constexpr int M = G::M;
  static_assert(F::D == G::D, "drift and diffusion disagree on d");
  // Column loops are unrolled for small systems (column index known at compile time: structural zeros
  // of G vanish).  SDB_SMALL_LIMIT lets the NVRTC path switch that off for user systems whose G is
  // expensive (every unrolled copy inlines the user's code again).
#ifndef SDB_SMALL_LIMIT
#define SDB_SMALL_LIMIT 36
#endif
  constexpr bool SMALL = (D * M <= SDB_SMALL_LIMIT);
  // Additive noise (G independent of y): the stage-2 differences g_k(H2) - g_k(H3) of SRK2 and the
  // G(Ybar) - G_n of KP2iS are exactly zero, so the repeated integrals are never used and are neither
  // generated nor read (bit-identical results; the first step of the reference's planned
  // noise-structure dispatch, integrate.py:150-151, README.rst:130).
  constexpr bool NEED_IJ =
      (SCHEME == SDB_SCHEME_SRK2 || SCHEME == SDB_SCHEME_KP2IS) && !G::ADDITIVE;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.P) return;

  double Y[D];
  if (a.flags & SDB_FLAG_Y0_BROADCAST) {
#pragma unroll
    for (int i = 0; i < D; ++i) Y[i] = a.y0[i];
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) Y[i] = a.y0[i * a.y0_sI + p * a.y0_sP];
  }
  sdb_store<D>(a, p, 0, Y);
#if SDB_OBSERVE
  SdbObs<D> obs;
  obs.init(a, Y);
#endif

  // KP2iS keeps the previous state, drift and noise term (integrate.py:621,628,635)
  double Ym1[SCHEME == SDB_SCHEME_KP2IS ? D : 1], fm1[SCHEME == SDB_SCHEME_KP2IS ? D : 1],
      Vm1[SCHEME == SDB_SCHEME_KP2IS ? D : 1];
  int64_t bad_solve = -1;

  const int steps = a.N - 1;
  int next_save = a.save_every;
  int64_t slot = 1;
#pragma unroll 1
  for (int n = 0; n < steps; ++n) {
    const double t0 = a.tspan[n];
    const double t1 = a.tspan[n + 1];
    double dW[M];
    double IJ[M][M];  // dead (and removed by the compiler) for Euler / Heun
    // Large systems (column loops rolled, G_n and I in local memory), SRI2 / SRS2 on device noise: keep
    // the Levy areas in registers and never build I (see the SRK2 branch)
    constexpr bool AREAS = !SMALL && NEED_IJ && SCHEME == SDB_SCHEME_SRK2 && !G::DIAGONAL &&
                           NOISE != SDB_NOISE_HOST && M >= 2 && M <= 12;
    double At[AREAS ? M * (M - 1) / 2 + 1 : 1];
    if constexpr (AREAS) {
      sdb_step_noise_areas<M, NOISE>(a, p, n, dW, At);
    } else {
      sdb_step_noise<M, NOISE, NEED_IJ, (G::DIAGONAL && SCHEME == SDB_SCHEME_SRK2)>(a, p, n, dW, IJ);
    }

    if (SCHEME == SDB_SCHEME_EULER) {
      const double h = a.h_global;
      double f0[D], Yn[D];
      F::eval(fp, Y, t0, f0);
#pragma unroll
      for (int i = 0; i < D; ++i) Yn[i] = fma(f0[i], h, Y[i]);
#pragma unroll(SMALL ? M : 1)
      for (int k = 0; k < M; ++k) {
        double g[D];
        G::col(gp, k, Y, t0, g);
#pragma unroll
        for (int i = 0; i < D; ++i) Yn[i] = fma(g[i], dW[k], Yn[i]);
      }
#pragma unroll
      for (int i = 0; i < D; ++i) Y[i] = Yn[i];
    } else if (SCHEME == SDB_SCHEME_HEUN) {
      const double h = a.h_global;
      double f0[D], Gw[D], Yb[D];
      F::eval(fp, Y, t0, f0);
#pragma unroll
      for (int i = 0; i < D; ++i) Gw[i] = 0.0;
#pragma unroll(SMALL ? M : 1)
      for (int k = 0; k < M; ++k) {
        double g[D];
        G::col(gp, k, Y, t0, g);
#pragma unroll
        for (int i = 0; i < D; ++i) Gw[i] = fma(g[i], dW[k], Gw[i]);
      }
#pragma unroll
      for (int i = 0; i < D; ++i) Yb[i] = fma(f0[i], h, Y[i]) + Gw[i];
      double f1[D], Gw1[D];
      F::eval(fp, Yb, t1, f1);
#pragma unroll
      for (int i = 0; i < D; ++i) Gw1[i] = 0.0;
#pragma unroll(SMALL ? M : 1)
      for (int k = 0; k < M; ++k) {
        double g[D];
        G::col(gp, k, Yb, t1, g);
#pragma unroll
        for (int i = 0; i < D; ++i) Gw1[i] = fma(g[i], dW[k], Gw1[i]);
      }
#pragma unroll
      for (int i = 0; i < D; ++i)
        Y[i] = Y[i] + 0.5 * (f0[i] + f1[i]) * h + 0.5 * (Gw[i] + Gw1[i]);
    } else if (SCHEME == SDB_SCHEME_R2) {
      // Burrage & Burrage's R2 (the reference's to-do list, README.rst:125-128): the same tableau for the
      // deterministic and the stochastic part,  Y2 = y + (2/3) (f h + G dW)(y),
      // y' = y + (1/4) (f h + G dW)(y) + (3/4) (f h + G dW)(Y2, t + 2h/3); strong order 1 for
      // Stratonovich equations with commutative noise (restated in oracle/sde_oracle.py: r2).
      const double h = a.h_global;
      const double tm = fma(2.0 / 3.0, h, t0);
      double f0[D], Gw[D], Yb[D];
      F::eval(fp, Y, t0, f0);
#pragma unroll
      for (int i = 0; i < D; ++i) Gw[i] = 0.0;
#pragma unroll(SMALL ? M : 1)
      for (int k = 0; k < M; ++k) {
        double g[D];
        G::col(gp, k, Y, t0, g);
#pragma unroll
        for (int i = 0; i < D; ++i) Gw[i] = fma(g[i], dW[k], Gw[i]);
      }
#pragma unroll
      for (int i = 0; i < D; ++i) Yb[i] = fma(2.0 / 3.0, fma(f0[i], h, Gw[i]), Y[i]);
      double f1[D], Gw1[D];
      F::eval(fp, Yb, tm, f1);
#pragma unroll
      for (int i = 0; i < D; ++i) Gw1[i] = 0.0;
#pragma unroll(SMALL ? M : 1)
      for (int k = 0; k < M; ++k) {
        double g[D];
        G::col(gp, k, Yb, tm, g);
#pragma unroll
        for (int i = 0; i < D; ++i) Gw1[i] = fma(g[i], dW[k], Gw1[i]);
      }
#pragma unroll
      for (int i = 0; i < D; ++i)
        Y[i] = Y[i] + 0.25 * fma(f0[i], h, Gw[i]) + 0.75 * fma(f1[i], h, Gw1[i]);
    } else if (SCHEME == SDB_SCHEME_SRK2) {
      const double h = t1 - t0;  // per-step h: integrate.py:497
      const double rh = sqrt(h);
      const double inv_rh = 1.0 / rh;
      double f0h[D], base[D], f1h[D], Yn[D];
      double Gn[D][M];
      F::eval(fp, Y, t0, f0h);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        f0h[i] *= h;
        base[i] = Y[i] + f0h[i];
      }
      F::eval(fp, base, t1, f1h);
#pragma unroll
      for (int i = 0; i < D; ++i) Yn[i] = Y[i] + 0.5 * fma(f1h[i], h, f0h[i]);
#pragma unroll(SMALL ? M : 1)
      for (int k = 0; k < M; ++k) {
        double g[D];
        G::col(gp, k, Y, t0, g);
#pragma unroll
        for (int i = 0; i < D; ++i) {
          Gn[i][k] = g[i];
          Yn[i] = fma(g[i], dW[k], Yn[i]);
        }
      }
      const double half_rh = 0.5 * rh;
      double S[AREAS ? D : 1][AREAS ? M : 1];
      if constexpr (AREAS) {
        // S = G_n I / sqrt(h) row by row, I = dW dW^T / 2 + diag + antisymmetric areas (sdb_assemble_IJ)
        // applied from registers: 2 local-memory accesses per entry of G_n instead of 2 m.
        const bool strat = (a.flags & SDB_FLAG_STRATONOVICH) != 0;
        const double dgv = strat ? 0.0 : -0.5 * a.h_global;
#pragma unroll 1
        for (int i = 0; i < D; ++i) {
          double gi[M], si[M];
          double gdw = 0.0;
#pragma unroll
          for (int j = 0; j < M; ++j) {
            gi[j] = Gn[i][j];
            gdw = fma(gi[j], dW[j], gdw);
          }
#pragma unroll
          for (int kk = 0; kk < M; ++kk) si[kk] = fma(0.5 * gdw, dW[kk], dgv * gi[kk]);
          int r = 0;
#pragma unroll
          for (int j = 0; j < M; ++j) {
#pragma unroll
            for (int kk = j + 1; kk < M; ++kk) {
              const double av = NOISE == SDB_NOISE_WIK ? -At[r] : At[r];
              si[kk] = fma(gi[j], av, si[kk]);
              si[j] = fma(-gi[kk], av, si[j]);
              ++r;
            }
          }
#pragma unroll
          for (int kk = 0; kk < M; ++kk) S[i][kk] = si[kk] * inv_rh;
        }
      }
#pragma unroll(SMALL ? M : 1)
      for (int k = 0; k < (NEED_IJ ? M : 0); ++k) {
        double up[D], dn[D], gu[D], gd[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
          double s;
          if constexpr (AREAS) {
            s = S[i][k];
          } else {
            s = Gn[i][0] * IJ[0][k];
#pragma unroll
            for (int j = 1; j < M; ++j) s = fma(Gn[i][j], IJ[j][k], s);
            s *= inv_rh;
          }
          up[i] = base[i] + s;
          dn[i] = base[i] - s;
        }
        G::col(gp, k, up, t1, gu);
        G::col(gp, k, dn, t1, gd);
#pragma unroll
        for (int i = 0; i < D; ++i) Yn[i] = fma(half_rh, gu[i] - gd[i], Yn[i]);
      }
#pragma unroll
      for (int i = 0; i < D; ++i) Y[i] = Yn[i];
    } else {  // KP2IS
      const double h = t1 - t0;
      const double rh = sqrt(h);
      double f0[D], V[D], acc[D];
      double Gn[D][M];
      F::eval(fp, Y, t0, f0);
#pragma unroll
      for (int i = 0; i < D; ++i) { V[i] = 0.0; acc[i] = 0.0; }
#pragma unroll(SMALL ? M : 1)
      for (int k = 0; k < M; ++k) {
        double g[D];
        G::col(gp, k, Y, t0, g);
#pragma unroll
        for (int i = 0; i < D; ++i) {
          Gn[i][k] = g[i];
          V[i] = fma(g[i], dW[k], V[i]);
        }
      }
      if constexpr (SdbGLinear<G>::value) {
        // Linear g_k = B_k y: G(Ybar_j1) - G_n = [B_k (f h + G_n[:, j1] sqrt h)]_k, so the m^2 column
        // evaluations of integrate.py:626-627 collapse to m:  sum1 = sum_k B_k v_k,
        // v_k = f h sum_j1 J[j1, k] + sqrt h (G_n J)[:, k]   (same value up to rounding)
#pragma unroll 1
        for (int k = 0; k < (NEED_IJ ? M : 0); ++k) {
          double v[D], g[D];
          double cs = 0.0;
#pragma unroll
          for (int j1 = 0; j1 < M; ++j1) cs += IJ[j1][k];
#pragma unroll
          for (int i = 0; i < D; ++i) {
            double s = Gn[i][0] * IJ[0][k];
#pragma unroll
            for (int j1 = 1; j1 < M; ++j1) s = fma(Gn[i][j1], IJ[j1][k], s);
            v[i] = fma(f0[i] * h, cs, rh * s);
          }
          G::col(gp, k, v, t0, g);
#pragma unroll
          for (int i = 0; i < D; ++i) acc[i] += g[i];
        }
      } else
#pragma unroll 1
      for (int j1 = 0; j1 < (NEED_IJ ? M : 0); ++j1) {
        double yb[D];
#pragma unroll
        for (int i = 0; i < D; ++i) yb[i] = fma(Gn[i][j1], rh, fma(f0[i], h, Y[i]));
#pragma unroll(SMALL ? M : 1)
        for (int k = 0; k < M; ++k) {
          double g[D];
          G::col(gp, k, yb, t0, g);  // G(Ybar[:,j1], tn): integrate.py:627
#pragma unroll
          for (int i = 0; i < D; ++i) acc[i] = fma(g[i] - Gn[i][k], IJ[j1][k], acc[i]);
        }
      }
      const double inv_rh = 1.0 / rh;
#pragma unroll
      for (int i = 0; i < D; ++i) V[i] = fma(acc[i], inv_rh, V[i]);
      double Ynew[D];
      if (n == 0) {
#pragma unroll
        for (int i = 0; i < D; ++i) Ynew[i] = fma(f0[i], h, Y[i]) + V[i];  // integrate.py:632
      } else {
        double cst[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
          const double gam = a.kp[i], al1 = a.kp[D + i], al2 = a.kp[2 * D + i];
          cst[i] = (1.0 - gam) * Y[i] + gam * Ym1[i] +
                   ((gam * al1 + (1.0 - al2)) * f0[i] + gam * (1.0 - al1) * fm1[i]) * h + V[i] +
                   gam * Vm1[i];
        }
        // Newton on r(x) = cst + al2 h f(x, t1) - x with a forward-difference Jacobian; the
        // reference hands the same residual to MINPACK hybrd (integrate.py:603-609,638).
#pragma unroll
        for (int i = 0; i < D; ++i) Ynew[i] = Y[i];
        bool conv = false;
        if constexpr (F::LINEAR) {
          // Linear drift f = A y on a uniform grid: the implicit equation (1 - diag(al2) h A) x = cst
          // is the SAME linear system for every path and step, so the host inverts it once
          // (sdb_capi.cu: kp[3 D ..] = (1 - diag(al2) h A)^-1, Gauss-Jordan with partial pivoting) and a
          // step is one matrix-vector product, followed by one Newton correction with the same
          // inverse (residual of the first solve: removes the rounding of the inverse).
          const double* __restrict__ Mi = a.kp + 3 * D;
          double x0[D], fx[D];
#pragma unroll
          for (int i = 0; i < D; ++i) {
            double acc = Mi[i * D] * cst[0];
#pragma unroll
            for (int j = 1; j < D; ++j) acc = fma(Mi[i * D + j], cst[j], acc);
            x0[i] = acc;
          }
          F::eval(fp, x0, t1, fx);
          double r[D];
#pragma unroll
          for (int i = 0; i < D; ++i) r[i] = fma(a.kp[2 * D + i] * h, fx[i], cst[i]) - x0[i];
#pragma unroll
          for (int i = 0; i < D; ++i) {
            double acc = Mi[i * D] * r[0];
#pragma unroll
            for (int j = 1; j < D; ++j) acc = fma(Mi[i * D + j], r[j], acc);
            Ynew[i] = x0[i] + acc;
          }
          conv = true;
        }
#pragma unroll 1
        for (int it = 0; it < 50 && !conv; ++it) {
          double fx[D], r[D];
          double Jm[D][D];
          F::eval(fp, Ynew, t1, fx);
#pragma unroll
          for (int i = 0; i < D; ++i) r[i] = fma(a.kp[2 * D + i] * h, fx[i], cst[i]) - Ynew[i];
          if constexpr (F::LINEAR) {  // d r / d x = al2 h A - 1, exactly
#pragma unroll 1
            for (int j = 0; j < D; ++j)
#pragma unroll
              for (int i = 0; i < D; ++i)
                Jm[i][j] = a.kp[2 * D + i] * h * F::jac(fp, i, j) - (i == j ? 1.0 : 0.0);
          } else {
#pragma unroll 1
            for (int j = 0; j < D; ++j) {
              double xp[D], fxp[D];
              double step = 1.4901161193847656e-08 * fabs(Ynew[j]);
              if (step == 0.0) step = 1.4901161193847656e-08;
#pragma unroll
              for (int i = 0; i < D; ++i) xp[i] = Ynew[i];
              xp[j] += step;
              step = xp[j] - Ynew[j];
              F::eval(fp, xp, t1, fxp);
#pragma unroll
              for (int i = 0; i < D; ++i)
                Jm[i][j] = a.kp[2 * D + i] * h * (fxp[i] - fx[i]) / step - (i == j ? 1.0 : 0.0);
            }
          }
#pragma unroll
          for (int i = 0; i < D; ++i) r[i] = -r[i];
          if (!sdb_solve<D>(Jm, r)) break;
          double nd = 0.0, nx = 0.0;
#pragma unroll
          for (int i = 0; i < D; ++i) {
            Ynew[i] += r[i];
            nd = fma(r[i], r[i], nd);
            nx = fma(Ynew[i], Ynew[i], nx);
          }
          conv = (nd <= a.rtol * a.rtol * nx) || nd == 0.0;
        }
        if (!conv && bad_solve < 0) bad_solve = n;
      }
#pragma unroll
      for (int i = 0; i < D; ++i) {
        Ym1[i] = Y[i];
        fm1[i] = f0[i];
        Vm1[i] = V[i];
        Y[i] = Ynew[i];
      }
    }

#if SDB_OBSERVE
    obs.step(a, n, Y);
#endif
    if (n + 1 == next_save) {
      sdb_store<D>(a, p, slot, Y);
      ++slot;
      next_save += a.save_every;
    }
  }
#if SDB_OBSERVE
  obs.finish(a, p);
#endif
  sdb_finish<D>(a, p, Y);
  if (SCHEME == SDB_SCHEME_KP2IS && bad_solve >= 0 && a.status) {
    atomicAdd((unsigned long long*)&a.status[2], 1ull);
    atomicMin((long long*)&a.status[3], (long long)bad_solve);
  }
}

// SDB_LOOP_MINBLOCKS (NVRTC builds): minimum resident blocks per SM, i.e. a register cap.
#ifndef SDB_LOOP_MINBLOCKS
#define SDB_LOOP_MINBLOCKS 1
#endif
#define SDB_LOOP_KERNEL(NAME, SCHEME, F, G, NOISE)                                              \
  extern "C" __global__ void __launch_bounds__(SDB_BLOCK, SDB_LOOP_MINBLOCKS)                   \
      NAME(const __grid_constant__ SdbLoopArgs a, const __grid_constant__ F::Params fp,         \
           const __grid_constant__ G::Params gp) {                                              \
    sdb_loop_body<SCHEME, F, G, NOISE>(a, fp, gp);                                              \
  }

// ---- standalone noise kernels ---------------------------------------------------------------------------
// deltaW (wiener.py:52): out[step][comp][path] = sqrt(h) * z, one thread per (path, step).
struct SdbDeltaWArgs {
  int64_t steps, paths, path_id0;
  uint64_t seed;
  int m;
  double sqrt_h;
  double* out; int64_t sN, sJ, sP;
};

#ifdef SDB_DEFINE_STANDALONE
extern "C" __global__ void __launch_bounds__(256) sdb_k_delta_w(const SdbDeltaWArgs a) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.steps * a.paths) return;
  const int64_t n = idx / a.paths;
  const int64_t p = idx - n * a.paths;
  const SdbPhiloxKey key(a.seed);
  const SdbStream st(key, (uint64_t)(a.path_id0 + p), (uint32_t)n,
                     reinterpret_cast<const double2*>(sdb_logtab), a.m > 2);
  SdbNormalSeq s(st, 0u);
  double* o = a.out + n * a.sN + p * a.sP;
  for (int j = 0; j < a.m; ++j) o[j * a.sJ] = a.sqrt_h * s.next();
}
#endif

// Ikpw/Jkpw/Iwik/Jwik for every (path, step): one thread each.
template <int M>
SDB_DEV void sdb_repint_body(const SdbRepintArgs& a) {
  constexpr int MM = M * (M - 1) / 2;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.steps * a.paths) return;
  const int64_t n = idx / a.paths;
  const int64_t p = idx - n * a.paths;
  double dW[M];
  const double* w = a.dW + n * a.dW_sN + p * a.dW_sP;
#pragma unroll
  for (int j = 0; j < M; ++j) dW[j] = w[j * a.dW_sJ];
  double At[MM + 1];
  if (a.X != nullptr) {
    const SdbNormalsHost<M> ns{a.X + n * a.XY_sN + p * a.XY_sP, a.Y + n * a.XY_sN + p * a.XY_sP,
                               a.XY_sT, a.XY_sJ,
                               a.Gt ? a.Gt + n * a.Gt_sN + p * a.Gt_sP : nullptr, a.Gt_sR};
    sdb_levy_areas<M>(dW, a.h, a.n_terms, a.method, ns, At);
  } else {
    const SdbPhiloxKey key(a.seed);
    const SdbStream st(key, (uint64_t)(a.path_id0 + p), (uint32_t)n);
    const SdbNormalsPhilox<M> ns{st, a.n_terms};
    sdb_levy_areas<M>(dW, a.h, a.n_terms, a.method, ns, At);
  }
  if (a.A != nullptr) {
    double* o = a.A + n * a.A_sN + p * a.A_sP;
    if (a.method == SDB_NOISE_KPW) {  // full antisymmetric m x m matrix A
      int r = 0;
#pragma unroll 1
      for (int i = 0; i < M; ++i) {
        o[(i * M + i) * a.A_sE] = 0.0;
#pragma unroll 1
        for (int j = i + 1; j < M; ++j) {
          o[(i * M + j) * a.A_sE] = At[r];
          o[(j * M + i) * a.A_sE] = -At[r];
          ++r;
        }
      }
    } else {  // Atilde, length M(M-1)/2 (at least one entry: zeros((N,1,1)) for m == 1)
#pragma unroll 1
      for (int r = 0; r < (MM > 0 ? MM : 1); ++r) o[r * a.A_sE] = At[r];
    }
  }
  // I (or J) assembled entry by entry from dW and the areas (no m x m temporary): KPW places +A
  // above the diagonal (wiener.py:106), Wiktorsson's column-major unvec below (wiener.py:278-280).
  double* q = a.IJ + n * a.IJ_sN + p * a.IJ_sP;
  const bool strat = a.stratonovich != 0;
  const double sg = a.method == SDB_NOISE_WIK ? -1.0 : 1.0;
  int r = 0;
#pragma unroll 1
  for (int i = 0; i < M; ++i) {
    double dii;
    if (M == 1 && a.method == SDB_NOISE_WIK)  // Iwik shortcut, wiener.py:259-260
      dii = strat ? (dW[i] * dW[i] - a.h) / 2.0 + 0.5 * a.h : (dW[i] * dW[i] - a.h) / 2.0;
    else
      dii = strat ? 0.5 * dW[i] * dW[i] : 0.5 * (dW[i] * dW[i] - a.h);
    q[i * a.IJ_sR + i * a.IJ_sC] = dii;
#pragma unroll 1
    for (int j = i + 1; j < M; ++j) {
      const double sym = 0.5 * dW[i] * dW[j];
      const double av = sg * At[r];
      q[i * a.IJ_sR + j * a.IJ_sC] = sym + av;
      q[j * a.IJ_sR + i * a.IJ_sC] = sym - av;
      ++r;
    }
  }
}

#define SDB_REPINT_KERNEL(NAME, M)                                                   \
  extern "C" __global__ void __launch_bounds__(128) NAME(const SdbRepintArgs a) {    \
    sdb_repint_body<M>(a);                                                           \
  }
