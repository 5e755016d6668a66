// sdb_fused_rt.h -- run-time mirror of SdbFusedShape / sdb_fused_build_tables (sdb_fused.cuh, default
// tuning: 4-row blocks, 256 threads) for shapes that have no ahead-of-time instantiation: the fused
// kernel is then compiled by NVRTC for the (d, m) at hand and this header sizes its shared memory
// and builds its DMMA fragment tables.  sdb_fused.cu static_asserts that both agree.
#pragma once
#include <stddef.h>

struct SdbFusedDims {
  int D, M, RB;  // RB: state rows per block (4: sdb_fused.cuh default; 2: sdb_fusedx.cuh)
  int XR;        // extra P1 row groups after the drift rows (1: the rows of A^2; -1: no drift rows, sdb_fused.cuh)
  constexpr SdbFusedDims(int d, int m, int rb = 4, int xr = 0) : D(d), M(m), RB(d > rb ? rb : 4), XR(xr) {}
  static constexpr int THREADS = 256, WARPS = 8, MAX_TERMS = 256;
  constexpr int MM() const { return M * (M - 1) / 2; }
  constexpr int NB() const { return (D + RB - 1) / RB; }
  constexpr int LT() const { return (D + 3) / 4; }
  constexpr int IT() const { return D / 8; }
  constexpr int IL() const { return D - 8 * IT(); }
  constexpr int rn(int b) const { return b < NB() - 1 ? RB : D - RB * (NB() - 1); }
  constexpr int r0(int b) const { return RB * b; }
  constexpr int nr(int b) const { return (M + 1 + XR) * rn(b); }
  constexpr int t1(int b) const { return (nr(b) + 7) / 8; }
  constexpr int kt3(int b) const { return (rn(b) * M + 3) / 4; }
  static constexpr int max2(int a, int b) { return a > b ? a : b; }
  constexpr int grows() const {
    int g = max2(8, D);
    for (int b = 0; b < NB(); ++b) g = max2(g, max2(8 * t1(b), 4 * kt3(b)));
    return g;
  }
  constexpr int p1_off(int b) const {
    int o = 0;
    for (int q = 0; q < b; ++q) o += t1(q) * LT();
    return o;
  }
  constexpr int p3_off(int b) const {
    int o = p1_off(NB());
    for (int q = 0; q < b; ++q) o += kt3(q) * IT();
    return o;
  }
  constexpr int n_frags() const { return p3_off(NB()); }
  constexpr int left_off(int b) const {
    int o = 0;
    for (int q = 0; q < b; ++q) o += IL() * rn(q) * M;
    return o;
  }
  constexpr int n_left() const { return left_off(NB()); }
  constexpr size_t table_doubles() const { return (size_t)n_frags() * 32 + n_left(); }
  constexpr size_t smem_base() const {
    return ((size_t)MM() * THREADS + (size_t)(grows() + M) * 32 * WARPS) * sizeof(double) +
           (MAX_TERMS + 1 + n_left() + 1) * sizeof(double);
  }
  constexpr int tab_copies() const { return smem_base() + 8 * 128 * 16 <= 227u * 1024u ? 8 : 1; }
  constexpr size_t smem_bytes() const { return smem_base() + (size_t)tab_copies() * 128 * 16; }
  // the kernel needs d < 16 (one DMMA tile of the stage-2 sum) and must fit an SM
  constexpr bool supported() const {
    return D >= 1 && M >= 2 && D < 16 && smem_bytes() <= 227u * 1024u;
  }
  // lane-major fragment tables: out[frag * 32 + lane], then the leftover rows (as sdb_fused_build_tables)
  void build_tables(const double* A, const double* B, double* out) const {
    double* A2 = new double[(size_t)D * D];
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) {
        double acc = 0.0;
        for (int l = 0; l < D && XR >= 0; ++l) acc += A[i * D + l] * A[l * D + j];
        A2[i * D + j] = acc;
      }
    for (int b = 0; b < NB(); ++b) {
      const int rnb = rn(b), r0b = r0(b);
      for (int t = 0; t < t1(b); ++t)
        for (int lt = 0; lt < LT(); ++lt)
          for (int lane = 0; lane < 32; ++lane) {
            const int n = 8 * t + (lane >> 2), l = 4 * lt + (lane & 3);
            double v = 0.0;
            if (l < D) {
              if (n < M * rnb) v = B[((n / rnb) * D + r0b + n % rnb) * D + l];
              else if (XR < 0) v = 0.0;  // any drift, evaluated per lane (SDB_FUSED_A2 -1): no drift rows
              else if (n < (M + 1) * rnb) v = A[(r0b + n - M * rnb) * D + l];
              else if (n < nr(b)) v = A2[(r0b + n - (M + 1) * rnb) * D + l];
            }
            out[(size_t)(p1_off(b) + t * LT() + lt) * 32 + lane] = v;
          }
      for (int kt = 0; kt < kt3(b); ++kt)
        for (int it = 0; it < IT(); ++it)
          for (int lane = 0; lane < 32; ++lane) {
            const int i = 8 * it + (lane >> 2), K = 4 * kt + (lane & 3);
            double v = 0.0;
            if (K < rnb * M) v = B[((K % M) * D + i) * D + r0b + K / M];
            out[(size_t)(p3_off(b) + kt * IT() + it) * 32 + lane] = v;
          }
    }
    double* left = out + (size_t)n_frags() * 32;
    for (int b = 0; b < NB(); ++b)
      for (int q = 0; q < IL(); ++q)
        for (int K = 0; K < rn(b) * M; ++K)
          left[left_off(b) + q * rn(b) * M + K] = B[((K % M) * D + 8 * IT() + q) * D + r0(b) + K / M];
    delete[] A2;
  }
};

// Run-time mirror of SdbXShape (sdb_fusedx.cuh): the TMEM-resident kernel, 2-row blocks.
struct SdbFusedXDims {
  SdbFusedDims sh;
  // 2-row blocks; xr = 1: rows of A^2 in the P1 product, -1: no drift rows (any drift, evaluated per lane:
  // D more shared-memory rows per warp hold f(Y_n) h)
  constexpr SdbFusedXDims(int d, int m, int xr = 1) : sh(d, m, 2, xr) {}
  constexpr int MM() const { return sh.MM(); }
  constexpr int MMP() const { return (MM() + 15) / 16 * 16; }
  constexpr int C_END() const { return 2 * MMP() + (2 * sh.M + 7) / 8 * 8 + 4 * sh.D; }
  constexpr int WPQ() const { return 2 * C_END() <= 512 ? 2 : 1; }
  constexpr int threads() const { return 128 * WPQ(); }
  constexpr size_t smem_bytes() const {
    const size_t b = (size_t)(sh.grows() + sh.D + (sh.XR < 0 ? sh.D : 0)) * 32 * 8 * 4 * WPQ() + 128 * 16 +
                     (SdbFusedDims::MAX_TERMS + 1 + sh.n_left() + 1) * sizeof(double) + 16;
    return b < 120u * 1024u ? 120u * 1024u : b;
  }
  // even m; uniform 2-row blocks (even d > 2); the areas, dW and row pieces of a path within the 512
  // columns of a TMEM lane; at most 3 DMMA row tiles of the stage-2 sum; one SM's shared memory
  constexpr bool supported() const {
    return sh.M >= 2 && sh.M % 2 == 0 && sh.D > 2 && sh.D % 2 == 0 && sh.D <= 24 && C_END() <= 512 &&
           smem_bytes() <= 227u * 1024u;
  }
};

// Run-time mirror of SdbX2Shape (sdb_fusedx2.cuh): the TMEM kernel with two lanes per path, for
// shapes whose areas leave room for only one warp per sub-partition in SdbFusedXDims.
struct SdbFusedX2Dims {
  SdbFusedXDims x;
  constexpr SdbFusedX2Dims(int d, int m, int xr = 1) : x(d, m, xr) {}   // xr -1: any drift (D more rows per warp)
  static constexpr int threads() { return 256; }
  static constexpr int paths_per_block() { return 128; }
  constexpr int GR() const { return x.sh.grows() >= 2 * x.sh.M ? x.sh.grows() : 2 * x.sh.M; }
  constexpr size_t smem_bytes() const {
    const size_t b = ((size_t)GR() * 32 * 8 + (size_t)x.sh.D * 32 * 4 +
                      (x.sh.XR < 0 ? (size_t)x.sh.D * 32 * 8 : 0)) * 8 + 128 * 16 +
                     (SdbFusedDims::MAX_TERMS + 1 + x.sh.n_left() + 1) * sizeof(double) + 16;
    return b < 120u * 1024u ? 120u * 1024u : b;
  }
  // an even number of row blocks and at least two area chunks to split
  constexpr bool supported() const {
    return x.supported() && x.WPQ() == 1 && x.sh.NB() % 2 == 0 && x.MMP() >= 32 &&
           smem_bytes() <= 227u * 1024u;
  }
};

// Run-time mirror of SdbRxShape (sdb_repintx.cuh).
struct SdbRepintXDims {
  int M;
  constexpr explicit SdbRepintXDims(int m) : M(m) {}
  constexpr int C_END() const { return 2 * ((M * (M - 1) / 2 + 15) / 16 * 16) + (2 * M + 7) / 8 * 8; }
  constexpr int threads() const { return 2 * C_END() <= 512 ? 256 : 128; }
  constexpr size_t smem_bytes() const { return 120u * 1024u; }
  constexpr bool supported() const { return M >= 2 && M % 2 == 0 && C_END() <= 512; }
};

// Run-time mirror of SdbEhShape (sdb_fusedeh.cuh): fused Euler (12 warps per SM) / Heun (8 warps).
struct SdbFusedEhDims {
  SdbFusedDims sh;
  int heun;
  constexpr SdbFusedEhDims(int d, int m, int heun_, int xr = 0) : sh(d, m, 4, xr), heun(heun_) {}
  constexpr int threads() const { return heun ? 256 : 384; }
  constexpr int grows() const {
    int g = SdbFusedDims::max2(8, sh.D);
    for (int b = 0; b < sh.NB(); ++b) g = SdbFusedDims::max2(g, 8 * sh.t1(b));
    return g;
  }
  constexpr size_t smem_bytes() const { return (size_t)grows() * 32 * 8 * (threads() / 32) + 128 * 16; }
  constexpr bool supported() const {
    return sh.M >= 2 && sh.D >= 1 && sh.D <= 16 && smem_bytes() <= 227u * 1024u;
  }
};

// Run-time mirror of SdbFcShape / sdb_fusedc_build_tables (sdb_fusedc.cuh): SRI2/SRS2 on the
// commutative-noise integrals, two P1 products per step.
struct SdbFusedCDims {
  SdbFusedDims s1, s2;  // pseudo-system [B_1 .. B_m, A^2, S ; A] and the plain [B_1 .. B_m ; A]
  constexpr SdbFusedCDims(int d, int m) : s1(d, m + 2, 4), s2(d, m, 4) {}
  static constexpr int threads() { return 256; }
  constexpr int grows() const {
    int g = SdbFusedDims::max2(8, s1.D);
    for (int b = 0; b < s1.NB(); ++b) g = SdbFusedDims::max2(g, 8 * s1.t1(b));
    for (int b = 0; b < s2.NB(); ++b) g = SdbFusedDims::max2(g, 8 * s2.t1(b));
    return g;
  }
  constexpr size_t table1_doubles() const { return s1.table_doubles(); }
  constexpr size_t table_doubles() const { return s1.table_doubles() + s2.table_doubles(); }
  constexpr size_t smem_bytes() const { return (size_t)grows() * 32 * 8 * (threads() / 32) + 128 * 16; }
  constexpr bool supported() const {
    return s2.M >= 2 && s2.D >= 1 && s2.D <= 12 && smem_bytes() <= 227u * 1024u;
  }
  void build_tables(const double* A, const double* B, double* out) const {
    const int D = s2.D, M = s2.M;
    double* Bx = new double[(size_t)(M + 2) * D * D];
    for (int i = 0; i < M * D * D; ++i) Bx[i] = B[i];
    double* A2 = Bx + (size_t)M * D * D;
    double* S = A2 + D * D;
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) {
        double a2 = 0.0, s = 0.0;
        for (int l = 0; l < D; ++l) a2 += A[i * D + l] * A[l * D + j];
        for (int k = 0; k < M; ++k)
          for (int l = 0; l < D; ++l) s += B[(k * D + i) * D + l] * B[(k * D + l) * D + j];
        A2[i * D + j] = a2;
        S[i * D + j] = s;
      }
    s1.build_tables(A, Bx, out);
    s2.build_tables(A, B, out + table1_doubles());
    delete[] Bx;
  }
};
