/* sdb.h -- C ABI of libsdb.so, the B200 (sm_100a) ensemble SDE integrator.
 *
 * The reference (mattja/sdeint, pure Python) has no FFI: its only boundary is
 * the Python function API of sdeint/__init__.py:3-5.  This header is the
 * boundary a maintainer of the reference would bind with ctypes to replace the
 * hot path (the per-step loops of sdeint/integrate.py and the noise
 * construction of sdeint/wiener.py); INTEGRATION.md shows that binding and
 * sdeint_b200/_lib.py is the one this repository ships.
 *
 * Conventions
 *   - every entry point returns 0 on success, non-zero on error; the message is
 *     available from sdb_last_error() (thread-local);
 *   - pointers named d_* are DEVICE pointers, h_* are HOST pointers; the caller
 *     owns every buffer (the Python shim allocates them as torch tensors);
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous;
 *   - arrays are addressed through explicit element strides (in doubles) so the
 *     same kernels serve the reference's per-path (N, .) layout and the
 *     path-minor ensemble layout [time][component][path] used for coalescing;
 *   - no torch types, no C++ types.
 */
#ifndef SDB_H
#define SDB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDB_VERSION 100

/* ---- schemes: the integrators of sdeint/integrate.py ------------------------------- */
#define SDB_SCHEME_EULER 0 /* itoEuler    integrate.py:190-240 */
#define SDB_SCHEME_HEUN 1  /* stratHeun   integrate.py:243-303 */
#define SDB_SCHEME_SRK2 2  /* itoSRI2 / stratSRS2 = _Roessler2010_SRK2, integrate.py:436-523 */
#define SDB_SCHEME_KP2IS 3 /* stratKP2iS  integrate.py:526-646; with Ito integrals I (no SDB_FLAG_STRATONOVICH)
                            * the Ito version of the same two-step implicit scheme (README.rst:125-128) */
#define SDB_SCHEME_R2 4    /* Burrage & Burrage (1996) two-stage SRK "R2", Stratonovich, commutative noise
                            * (README.rst:125-128 to-do); needs only dW, like SDB_SCHEME_HEUN */

/* ---- drift / diffusion operator kinds (sdeint_b200/ops.py) ------------------------- */
#define SDB_DRIFT_LINEAR 0 /* params A[d][d]               f = A y */
#define SDB_DRIFT_KP4446 1 /* params a, c                  f = -(a + c y)(1 - y^2) */
#define SDB_DRIFT_KPS445 2 /* params s                     f = -sin 2y - sin 4y / 4 + 2 s sin y cos^3 y */
#define SDB_DRIFT_USER 9   /* NVRTC: __device__ void sde_drift(const double* y, double t, const double* p, double* out) */
#define SDB_DIFF_CONST 0       /* params B[d][m]           G = B */
#define SDB_DIFF_LINEAR_COLS 1 /* params B[m][d][d]        g_k = B_k y */
#define SDB_DIFF_KP4446 2      /* params b                 G = b (1 - y^2) */
#define SDB_DIFF_KPS445 3      /* params sqrt2             G = sqrt2 cos^2 y */
#define SDB_DIFF_DIAG_LINEAR 4 /* params b[d]              g_k = b_k y_k e_k, m = d: DIAGONAL noise */
#define SDB_DIFF_AFFINE_COLS 5 /* params B[m][d][d], C[d][m]  g_k = B_k y + c_k (multiplicative + additive noise) */
#define SDB_DIFF_USER 9 /* NVRTC: __device__ void sde_diffusion_col(int k, const double* y, double t, const double* p, double* out) */
#define SDB_DIFF_USER_DIAG 10 /* as SDB_DIFF_USER, and the caller guarantees diagonal noise (m = d, column k has
                                 only component k and depends only on y_k): SRI2/SRS2 then need only the
                                 diagonal I_kk = (dW_k^2 - h)/2, no Levy areas (README.rst:130 of the reference) */

/* ---- where the noise comes from ------------------------------------------------------ */
#define SDB_NOISE_HOST 0 /* dW (and I/J for SRK2, KP2iS) supplied by the caller (dW=, I=/J=) */
#define SDB_NOISE_KPW 1  /* on-device Philox + Ikpw/Jkpw   wiener.py:77-131 */
#define SDB_NOISE_WIK 2  /* on-device Philox + Iwik/Jwik   wiener.py:234-305 */
#define SDB_NOISE_COMM 3 /* on-device Philox dW only; I = (dW dW^T - h 1)/2, J = dW dW^T / 2 with no Levy
                            areas: COMMUTATIVE noise (L^j g_k = L^k g_j), where the antisymmetric part of
                            I/J cancels in the order-1.0 term -- exactly so in SRI2/SRS2 for linear g_k = B_k y
                            with commuting B_k.  The fast path the reference plans at integrate.py:150-151 */

/* ---- flags --------------------------------------------------------------------------- */
#define SDB_FLAG_Y0_BROADCAST 1 /* d_y0 holds one state (d doubles) used by every path */
#define SDB_FLAG_STRATONOVICH 2 /* device noise: build J = I + h/2 (Jkpw/Jwik) instead of I */

typedef struct sdb_system sdb_system; /* opaque: drift + diffusion + compiled kernels */

const char* sdb_last_error(void);
int sdb_version(void);

/* Number of kernels launched by this library in the calling process (bench "gpu_launches"). */
uint64_t sdb_launch_count(void);

/* NVRTC-built kernels (user drift / diffusion source, shapes without an ahead-of-time kernel) are kept as
 * cubins under $SDB_CACHE_DIR (default ~/.cache/sdeint_b200; SDB_JIT_CACHE=0 disables), keyed by a hash of
 * the generated source, the embedded headers, the options and the NVRTC / library versions: how many
 * kernels of this process came from that cache, and how many were compiled.  (The reference has no
 * counterpart: its f and G are Python callables, integrate.py:124-146.) */
void sdb_jit_cache_stats(long* hits, long* misses);

/* A (drift, diffusion) pair of built-in kinds; parameters are copied.
 * Replaces the callables f, G of every integrator (integrate.py:124-146). */
int sdb_system_builtin(int drift_kind, const double* h_drift_params, int64_t n_drift,
                       int diff_kind, const double* h_diff_params, int64_t n_diff,
                       int d, int m, sdb_system** out);

/* Same, with user CUDA source for the kinds given as *_USER (NVRTC, sm_100a). `src` must
 * define sde_drift and/or sde_diffusion_col (see the kind comments above). */
int sdb_system_nvrtc(const char* src, int drift_kind, const double* h_drift_params,
                     int64_t n_drift, int diff_kind, const double* h_diff_params,
                     int64_t n_diff, int d, int m, sdb_system** out);

int sdb_system_destroy(sdb_system* sys);

/* Ensemble integration arguments.  Strides are in elements (doubles). */
typedef struct sdb_integrate_args {
  int scheme;          /* SDB_SCHEME_* */
  int noise;           /* SDB_NOISE_* */
  int flags;           /* SDB_FLAG_* */
  int n_terms;         /* series terms n of Ikpw/Iwik (default 5, wiener.py:77) */
  int64_t paths;       /* P: sample paths in this call */
  int64_t path_id0;    /* global id of the first path (Philox counter; rank offset) */
  uint64_t seed;       /* Philox key */
  int N;               /* len(tspan) */
  int save_every;      /* store y at steps 0, save_every, 2 save_every, ... */
  const double* h_tspan; /* HOST [N], equally spaced (checked by the caller) */
  const double* d_y0;  /* [d] if Y0_BROADCAST else strided */
  int64_t y0_stride_comp, y0_stride_path;
  double* d_y;         /* output, n_saved = (N-1)/save_every + 1 points */
  int64_t y_stride_save, y_stride_comp, y_stride_path;
  const double* d_dW;  /* NOISE_HOST: [N-1][m][P] through strides; else NULL */
  int64_t dW_stride_step, dW_stride_comp, dW_stride_path;
  const double* d_IJ;  /* NOISE_HOST, SRK2/KP2IS: [N-1][m][m][P] through strides */
  int64_t IJ_stride_step, IJ_stride_row, IJ_stride_col, IJ_stride_path;
  const double* h_kp2is; /* KP2IS: HOST [3*d] = gam, al1, al2 (integrate.py:587-592) */
  double rtol;         /* KP2IS: tolerance of the implicit solve (integrate.py:527) */
  int64_t* d_status;   /* [4]: #non-finite paths, first such path, #unconverged solves, first such step */
  void* stream;
} sdb_integrate_args;

/* The per-step loops of integrate.py:235-239, :292-302, :494-522, :613-645 over P paths. */
int sdb_integrate(sdb_system* sys, const sdb_integrate_args* args);

/* ---- path functionals -------------------------------------------------------------------
 * Quantities of the full-resolution trajectory that a decimated store (save_every > 1) cannot give
 * back, evaluated at every grid point t_0..t_{N-1} inside the step loop (the first-passage / `nsim`
 * role of the reference's to-do list, README.rst:137).  Up to SDB_MAX_OBSERVERS per call. */
#define SDB_MAX_OBSERVERS 4
#define SDB_OBS_FIRST_ABOVE 0 /* first grid time with y[comp] >= level (+inf if never) */
#define SDB_OBS_FIRST_BELOW 1 /* first grid time with y[comp] <= level (+inf if never) */
#define SDB_OBS_MAX 2         /* max over the grid of y[comp] */
#define SDB_OBS_MIN 3         /* min over the grid of y[comp] */
#define SDB_OBS_INTEGRAL 4    /* trapezoidal integral of y[comp] dt over the grid */
#define SDB_OBS_LAGPROD 5     /* sum_{n >= L} y[comp](t_n) y[comp](t_{n-L}), L = (int)level in [0, 32] grid steps:
                               * time-lagged products at full resolution (autocovariance / spectrum of a
                               * stationary path, the check test_integrate.py:455-456 leaves as a to-do) */
#define SDB_OBS_MAXLAG 32     /* longest lag of SDB_OBS_LAGPROD, in grid steps */
typedef struct sdb_observer {
  int kind;     /* SDB_OBS_* */
  int comp;     /* state component, 0 <= comp < d */
  double level; /* threshold of the FIRST_* kinds; lag in grid steps of SDB_OBS_LAGPROD */
} sdb_observer;

/* sdb_integrate plus n_obs observers: d_obs[o * stride_obs + p * stride_path] receives observer o of
 * path p.  Same kernels and results as sdb_integrate; the observing variants are compiled with NVRTC
 * on first use per system. */
int sdb_integrate_observed(sdb_system* sys, const sdb_integrate_args* args, const sdb_observer* h_obs,
                           int n_obs, double* d_obs, int64_t stride_obs, int64_t stride_path);

/* ---- resumed / segmented runs --------------------------------------------------------------
 * The on-device noise of interval n of path p is a pure function of (seed, path id, step counter), so a
 * run can be cut anywhere in time and continued later -- checkpoint / resume, or full-resolution output
 * produced segment by segment -- with bit-identical results: integrate the second segment with the
 * states the first one reached (d_y0 per path) and step0 = number of intervals already done.
 * h_noise > 0 fixes the variance of the increments to that of the whole run (the h of integrate.py:228,
 * :480; a segment's own (t[N-1]-t[0])/(N-1) may differ from it in the last bits).  Not for
 * SDB_SCHEME_KP2IS, a two-step scheme whose first step is special (integrate.py:632).  Observers as in
 * sdb_integrate_observed (n_obs = 0: none); they restart in every segment. */
typedef struct sdb_integrate_ext {
  int64_t step0;
  double h_noise;
  int n_obs;
  const sdb_observer* h_obs;
  double* d_obs;
  int64_t obs_stride_obs, obs_stride_path;
} sdb_integrate_ext;
int sdb_integrate_segment(sdb_system* sys, const sdb_integrate_args* args, const sdb_integrate_ext* ext);

/* Name of the kernel the last sdb_integrate* call of this thread launched: an ahead-of-time kernel's
 * symbol (e.g. "sdbk_fused_srk2_d10_m10_v0") or "jit:<cache key>" for an NVRTC-built one.  With the
 * environment variable SDB_TRACE set, every dispatch is also logged to stderr. */
const char* sdb_last_kernel(void);

/* deltaW (wiener.py:36-52) for P paths: out[step][comp][path] through strides. */
int sdb_delta_w(int64_t steps, int m, double h, int64_t paths, int64_t path_id0, uint64_t seed,
                double* d_out, int64_t stride_step, int64_t stride_comp, int64_t stride_path,
                void* stream);

/* Ikpw/Jkpw/Iwik/Jwik (wiener.py:77-131, :234-305) for P paths x `steps` intervals.
 * method: SDB_NOISE_KPW, SDB_NOISE_WIK or SDB_NOISE_COMM (A = 0, no normals drawn, n_terms ignored);
 * stratonovich: 0 -> I, 1 -> J.
 * d_X, d_Y ([n][step][comp][path]) and d_Gt ([step][pair][path]) are OPTIONAL caller-supplied
 * standard normals (construction parity with the reference's draws); when NULL they come
 * from Philox(seed, path_id0 + p, step).  d_A receives A (KPW: m x m) or Atilde (WIK: M). */
typedef struct sdb_repint_args {
  int method, stratonovich, n_terms, m;
  int64_t steps, paths, path_id0;
  uint64_t seed;
  double h;
  const double* d_dW; int64_t dW_stride_step, dW_stride_comp, dW_stride_path;
  const double* d_X;  const double* d_Y;
  int64_t XY_stride_term, XY_stride_step, XY_stride_comp, XY_stride_path;
  const double* d_Gt; int64_t Gt_stride_step, Gt_stride_pair, Gt_stride_path;
  double* d_A;  int64_t A_stride_step, A_stride_elem, A_stride_path;
  double* d_IJ; int64_t IJ_stride_step, IJ_stride_row, IJ_stride_col, IJ_stride_path;
  void* stream;
} sdb_repint_args;
int sdb_repeated_integrals(const sdb_repint_args* args);

/* Ensemble moments and histograms of saved states y[save][comp][path] (path stride 1):
 * d_sums[save][comp][2] += (sum y, sum y^2) in a fixed, partition-independent order per
 * block of 4096 paths; d_hist[save][comp][bins] += counts on [lo, hi) (NULL to skip). */
int sdb_moments(const double* d_y, int64_t n_saved, int d, int64_t paths, double* d_sums,
                unsigned long long* d_hist, int bins, double lo, double hi, void* stream);

/* Ensemble (count, mean, M2 = sum (y - mean)^2) per (saved point, component), the numerically safe
 * form of the variance: every chunk of 4096 paths is centred on its own mean and the
 * chunks are merged in chunk order with Chan's update; d_out[save][comp][3] is UPDATED by the same
 * merge (zero it for a fresh result), so successive calls -- or the per-rank results gathered by
 * the caller, merged in rank order -- accumulate deterministically. */
int sdb_moments_welford(const double* d_y, int64_t n_saved, int d, int64_t paths, double* d_out,
                        void* stream);

/* Ensemble second moments of saved states y[save][comp][path] (path stride 1):
 * d_out[save][i][j] += sum_p y_i y_j (full symmetric d x d matrix, d <= 16), summed per chunk of
 * 4096 paths and then over chunks in a fixed order (deterministic, partition-independent per chunk).
 * With sdb_moments this gives the ensemble covariance at every saved time; like the moment sums it
 * is all-reduced over ranks by the caller. */
int sdb_cross_moments(const double* d_y, int64_t n_saved, int d, int64_t paths, double* d_out,
                      void* stream);

/* Strided device -> host copy on `stream` (cudaMemcpy2DAsync): `height` rows of `width_bytes`,
 * row pitches in bytes.  The shim uses it to stream chunks of a path-minor trajectory buffer
 * into the caller's (pinned) result array while the next chunk of paths is being integrated. */
int sdb_copy2d_d2h_async(void* h_dst, int64_t dst_pitch_bytes, const void* d_src,
                         int64_t src_pitch_bytes, int64_t width_bytes, int64_t height, void* stream);

/* Measured peak of the FP64 pipe: runs a register-resident DFMA chain on every SM and
 * returns TFLOP/s (roofline denominator of bench.py). */
int sdb_fp64_peak(int iters, double* tflops_out, void* stream);

/* Raw generator access (tests): the Philox4x32-10 block with counter (path lo, path hi, step, slot)
 * and key seed (Random123 known-answer vectors), and the 128-bit block the kernels turn into the
 * Box-Muller pair `slot` of (seed, path, step): slot 0 is that Philox block at counter word 3 = 0,
 * slot s >= 1 is the MX2 expansion MX2(K, s) of the key block K = Philox at counter word
 * 3 = 0xFFFFFFFF (csrc/sdb_kernels.cuh; restated in oracle/philox_oracle.py).  Replaces numpy's
 * Generator.normal / standard_normal streams of sdeint/wiener.py:52,70-71,275. */
int sdb_philox_block(uint64_t seed, uint64_t path, uint32_t step, uint32_t slot, uint32_t out[4]);
int sdb_random_block(uint64_t seed, uint64_t path, uint32_t step, uint32_t slot, uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* SDB_H */
