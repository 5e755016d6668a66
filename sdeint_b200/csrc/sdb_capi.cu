// sdb_capi.cu -- the C ABI of libsdb.so (include/sdb.h): system objects, kernel lookup
// (ahead-of-time table, else NVRTC), launches.  No CPU fallback: every entry point that
// computes launches a CUDA kernel or fails with an error.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#include <atomic>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#define SDB_DEFINE_STANDALONE 1
#include "../../include/sdb.h"
#include "sdb_host.h"
#include "sdb_kernels.cuh"
#include "sdb_fused.cuh"
#include "sdb_fused_rt.h"

#include <math.h>
#include <stdlib.h>

// ---------------------------------------------------------------------------------------------
// errors, counters
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local std::string g_last_kernel;  // which kernel the last sdb_integrate of this thread launched
static std::atomic<uint64_t> g_launches{0};

static int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}
#define SDB_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess)                                                             \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e__));                 \
  } while (0)

extern "C" const char* sdb_last_error(void) { return g_err.c_str(); }
extern "C" int sdb_version(void) { return SDB_VERSION; }
extern "C" uint64_t sdb_launch_count(void) { return g_launches.load(); }

std::vector<SdbAotEntry>& sdb_aot_registry() {
  static std::vector<SdbAotEntry> reg;
  return reg;
}

std::string sdb_loop_key(int scheme, int fk, int gk, int d, int m, int noise, bool inl) {
  char buf[128];
  snprintf(buf, sizeof buf, "loop|s%d|f%d|g%d|d%d|m%d|n%d|p%c", scheme, fk, gk, d, m, noise,
           inl ? 'i' : 'g');
  return buf;
}

static const SdbAotEntry* aot_entry(const std::string& key) {
  for (const auto& e : sdb_aot_registry())
    if (e.key == key) return &e;
  return nullptr;
}

static const void* aot_lookup(const std::string& key) {
  const SdbAotEntry* e = aot_entry(key);
  return e ? e->fn : nullptr;
}

// The small stream-ordered scratch buffers (tspan / h tables, moment partials) come from the device's
// default memory pool.  Its default release threshold is 0, i.e. the pool hands everything back to
// the driver at every synchronisation and the next cudaMallocAsync pays for a fresh allocation;
// keep the pool's memory instead (set once per device).
static void keep_pool_memory() {
  static std::mutex mu;
  static bool done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  std::lock_guard<std::mutex> lk(mu);
  if (done[dev]) return;
  done[dev] = true;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
}

// Blocks for `total` work items, or 0 when the count does not fit a 1-d grid (2^31 - 1 blocks).
static unsigned blocks_for(int64_t total, int64_t per_block) {
  const int64_t b = (total + per_block - 1) / per_block;
  return b > 0x7fffffffll ? 0u : (unsigned)b;
}

static int launch(const void* fn, dim3 grid, dim3 block, void** args, cudaStream_t st,
                  size_t smem = 0) {
  if (grid.x == 0) return fail("launch: too many work items for one grid; split the call");
  SDB_CUDA(cudaLaunchKernel(fn, grid, block, args, smem, st));
  g_launches.fetch_add(1);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// NVRTC (loaded lazily with dlopen so that libsdb.so itself loads on machines without it)
// ---------------------------------------------------------------------------------------------
#include "sdb_kernels_embed.inc"  // kKernelHeader, kTablesHeader

typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* lib = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*,
                       const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*Version)(int*, int*) = nullptr;
};

static int nvrtc_load(Nvrtc** out) {
  static Nvrtc rt;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (!rt.lib) {
    const char* env = getenv("SDB_NVRTC_PATH");
    const char* names[] = {env, "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "libnvrtc.so"};
    for (const char* n : names) {
      if (!n) continue;
      rt.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (rt.lib) break;
    }
    if (!rt.lib) return fail("NVRTC (libnvrtc.so.12) not found; set SDB_NVRTC_PATH");
#define SDB_SYM(name)                                                              \
  *(void**)(&rt.name) = dlsym(rt.lib, "nvrtc" #name);                              \
  if (!rt.name) return fail("NVRTC symbol nvrtc" #name " missing");
    SDB_SYM(CreateProgram) SDB_SYM(CompileProgram) SDB_SYM(GetProgramLogSize)
    SDB_SYM(GetProgramLog) SDB_SYM(GetCUBINSize) SDB_SYM(GetCUBIN) SDB_SYM(DestroyProgram)
    SDB_SYM(GetErrorString) SDB_SYM(Version)
#undef SDB_SYM
  }
  *out = &rt;
  return 0;
}

struct JitKernel {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kern = nullptr;
};

// Bytes of machine code in a cubin: the sum of its .text.* sections (ELF64; -lineinfo makes the file
// size useless as a measure).
static size_t cubin_text_bytes(const std::vector<char>& elf) {
  struct Ehdr { unsigned char ident[16]; uint16_t type, machine; uint32_t version; uint64_t entry, phoff, shoff;
                uint32_t flags; uint16_t ehsize, phentsize, phnum, shentsize, shnum, shstrndx; };
  struct Shdr { uint32_t name, type; uint64_t flags, addr, offset, size; uint32_t link, info;
                uint64_t addralign, entsize; };
  if (elf.size() < sizeof(Ehdr)) return 0;
  Ehdr eh;
  memcpy(&eh, elf.data(), sizeof eh);
  if (memcmp(eh.ident, "\177ELF", 4) != 0 || eh.shentsize != sizeof(Shdr) ||
      eh.shoff + (uint64_t)eh.shnum * sizeof(Shdr) > elf.size() || eh.shstrndx >= eh.shnum)
    return 0;
  std::vector<Shdr> sh(eh.shnum);
  memcpy(sh.data(), elf.data() + eh.shoff, sh.size() * sizeof(Shdr));
  const Shdr& str = sh[eh.shstrndx];
  if (str.offset + str.size > elf.size()) return 0;
  size_t total = 0;
  for (const Shdr& x : sh) {
    if (x.name >= str.size) continue;
    const char* nm = elf.data() + str.offset + x.name;
    if (strncmp(nm, ".text.", 6) == 0) total += (size_t)x.size;
  }
  return total;
}

// ---- on-disk cache of NVRTC results ------------------------------------------------------------
// A fused kernel takes NVRTC several seconds; a user system is compiled again by every process that
// integrates it.  The cubin is therefore kept under $SDB_CACHE_DIR (default $XDG_CACHE_HOME/sdeint_b200 or
// ~/.cache/sdeint_b200), named by a 128-bit hash of everything that determines it: the generated source
// (which contains the user's code and every #define), the embedded headers, the compile options, the NVRTC
// version and this library's version.  SDB_JIT_CACHE=0 switches it off; an unwritable directory is ignored.
static uint64_t fnv1a64(const void* data, size_t n, uint64_t h) {
  const unsigned char* p = (const unsigned char*)data;
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001B3ull; }
  return h;
}
static std::string jit_cache_dir() {
  const char* off = getenv("SDB_JIT_CACHE");
  if (off && !strcmp(off, "0")) return "";
  if (const char* d = getenv("SDB_CACHE_DIR")) return d;
  if (const char* x = getenv("XDG_CACHE_HOME")) if (*x) return std::string(x) + "/sdeint_b200";
  if (const char* h = getenv("HOME")) if (*h) return std::string(h) + "/.cache/sdeint_b200";
  return "";
}
static void mkdirs(const std::string& dir) {
  for (size_t i = 1; i <= dir.size(); ++i)
    if (i == dir.size() || dir[i] == '/') mkdir(dir.substr(0, i).c_str(), 0755);
}
static std::atomic<long> g_jit_hits{0}, g_jit_misses{0};
extern "C" void sdb_jit_cache_stats(long* hits, long* misses) {
  if (hits) *hits = g_jit_hits.load();
  if (misses) *misses = g_jit_misses.load();
}

static int jit_compile(const std::string& source, const char* kernel_name, JitKernel* out,
                       size_t* text_bytes = nullptr) {
  Nvrtc* rt;
  if (nvrtc_load(&rt)) return 1;
  nvrtcProgram prog;
  const char* hdr_src[] = {kKernelHeader, kTablesHeader, kFusedHeader, kTmemHeader, kFusedWsHeader,
                           kFusedXHeader, kRepintXHeader, kFusedEhHeader, kFusedCHeader, kFusedX2Header};
  static const char* kOpts = "--gpu-architecture=sm_100a -std=c++17 -lineinfo --fmad=true -default-device";
  std::string cache_file;
  const std::string dir = jit_cache_dir();
  // cubins of this process by the same key: a sweep that changes only parameter VALUES creates a new system
  // per point but the same source -- it is compiled (or read from disk) once
  static std::map<std::pair<uint64_t, uint64_t>, std::vector<char>> mem_cache;
  static std::mutex mem_mu;
  std::pair<uint64_t, uint64_t> mem_key(0, 0);
  {
    static uint64_t hdr_hash[2] = {0, 0};
    static std::once_flag once;
    std::call_once(once, [&] {
      uint64_t a = 0xCBF29CE484222325ull, b = 0x9AE16A3B2F90404Full;
      for (const char* h : hdr_src) { a = fnv1a64(h, strlen(h), a); b = fnv1a64(h, strlen(h), b); }
      int major = 0, minor = 0;
      if (rt->Version) rt->Version(&major, &minor);
      const int ver[3] = {major, minor, SDB_VERSION};
      a = fnv1a64(ver, sizeof ver, a); b = fnv1a64(ver, sizeof ver, b);
      a = fnv1a64(kOpts, strlen(kOpts), a); b = fnv1a64(kOpts, strlen(kOpts), b);
      hdr_hash[0] = a; hdr_hash[1] = b;
    });
    const uint64_t a = fnv1a64(source.data(), source.size(), hdr_hash[0]);
    const uint64_t b = fnv1a64(source.data(), source.size(), hdr_hash[1]);
    mem_key = std::make_pair(a, b);
    {
      std::lock_guard<std::mutex> lk(mem_mu);
      auto it = mem_cache.find(mem_key);
      if (it != mem_cache.end() &&
          cudaLibraryLoadData(&out->lib, it->second.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) == cudaSuccess &&
          cudaLibraryGetKernel(&out->kern, out->lib, kernel_name) == cudaSuccess) {
        if (text_bytes) *text_bytes = cubin_text_bytes(it->second);
        g_jit_hits.fetch_add(1);
        return 0;
      }
    }
    char name[64];
    snprintf(name, sizeof name, "/%016llx%016llx.cubin", (unsigned long long)a, (unsigned long long)b);
    if (!dir.empty()) cache_file = dir + name;
    if (FILE* fh = cache_file.empty() ? nullptr : fopen(cache_file.c_str(), "rb")) {
      std::vector<char> cubin;
      char buf[1 << 16];
      size_t n;
      while ((n = fread(buf, 1, sizeof buf, fh)) > 0) cubin.insert(cubin.end(), buf, buf + n);
      fclose(fh);
      if (cubin.size() > 64 &&
          cudaLibraryLoadData(&out->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) == cudaSuccess &&
          cudaLibraryGetKernel(&out->kern, out->lib, kernel_name) == cudaSuccess) {
        if (text_bytes) *text_bytes = cubin_text_bytes(cubin);
        g_jit_hits.fetch_add(1);
        std::lock_guard<std::mutex> lk(mem_mu);
        if (mem_cache.size() >= 64) mem_cache.clear();
        mem_cache[mem_key] = std::move(cubin);
        return 0;
      }
      cudaGetLastError();  // a truncated or foreign file: compile as if it were not there
    }
  }
  g_jit_misses.fetch_add(1);
  const char* hdr_name[] = {"sdb_kernels.cuh", "sdb_tables.cuh", "sdb_fused.cuh", "sdb_tmem.cuh",
                            "sdb_fusedws.cuh", "sdb_fusedx.cuh", "sdb_repintx.cuh", "sdb_fusedeh.cuh",
                            "sdb_fusedc.cuh", "sdb_fusedx2.cuh"};
  int rc = rt->CreateProgram(&prog, source.c_str(), "sdb_jit.cu", 10, hdr_src, hdr_name);
  if (rc) return fail(std::string("nvrtcCreateProgram: ") + rt->GetErrorString(rc));
  // -default-device: un-annotated functions (the generic lambdas of the compile-time loops) are
  // device functions in JIT mode
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo",
                        "--fmad=true", "-default-device"};  // (kOpts above is the same list, hashed)
  rc = rt->CompileProgram(prog, 5, opts);
  if (rc) {
    size_t n = 0;
    rt->GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) rt->GetProgramLog(prog, &log[0]);
    rt->DestroyProgram(&prog);
    return fail("NVRTC compile failed:\n" + log);
  }
  size_t sz = 0;
  rt->GetCUBINSize(prog, &sz);
  std::vector<char> cubin(sz);
  rt->GetCUBIN(prog, cubin.data());
  rt->DestroyProgram(&prog);
  if (text_bytes) *text_bytes = cubin_text_bytes(cubin);
  SDB_CUDA(cudaLibraryLoadData(&out->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  SDB_CUDA(cudaLibraryGetKernel(&out->kern, out->lib, kernel_name));
  {
    std::lock_guard<std::mutex> lk(mem_mu);
    if (mem_cache.size() >= 64) mem_cache.clear();
    mem_cache[mem_key] = cubin;
  }
  if (!cache_file.empty()) {  // publish atomically: write a private file, then rename
    mkdirs(dir);
    const std::string tmp = cache_file + "." + std::to_string((long)getpid()) + ".tmp";
    if (FILE* fh = fopen(tmp.c_str(), "wb")) {
      const bool ok = fwrite(cubin.data(), 1, cubin.size(), fh) == cubin.size();
      if (fclose(fh) == 0 && ok) rename(tmp.c_str(), cache_file.c_str());
      else remove(tmp.c_str());
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// systems
// ---------------------------------------------------------------------------------------------
struct sdb_system {
  int fk, gk, d, m;
  std::vector<double> fp, gp;
  std::string src;           // user CUDA source (may be empty)
  double* d_fp = nullptr;    // device copies (PGlobal mode / user operators)
  double* d_gp = nullptr;
  std::map<const void*, double*> d_frags;  // fused kernels: DMMA fragment tables, per kernel
  bool inline_params = true;
  std::map<std::string, JitKernel> jit;
  std::mutex mu;
};

static const size_t kParamSpace = 32764;

static int expected_params(int kind, bool drift, int d, int m, int64_t* n) {
  if (drift) {
    switch (kind) {
      case SDB_DRIFT_LINEAR: *n = (int64_t)d * d; return 0;
      case SDB_DRIFT_KP4446: *n = 2; return d == 1 ? 0 : 1;
      case SDB_DRIFT_KPS445: *n = 1; return d == 1 ? 0 : 1;
      case SDB_DRIFT_USER: *n = -1; return 0;
    }
  } else {
    switch (kind) {
      case SDB_DIFF_CONST: *n = (int64_t)d * m; return 0;
      case SDB_DIFF_LINEAR_COLS: *n = (int64_t)m * d * d; return 0;
      case SDB_DIFF_KP4446: *n = 1; return (d == 1 && m == 1) ? 0 : 1;
      case SDB_DIFF_KPS445: *n = 1; return (d == 1 && m == 1) ? 0 : 1;
      case SDB_DIFF_DIAG_LINEAR: *n = d; return d == m ? 0 : 1;
      case SDB_DIFF_AFFINE_COLS: *n = (int64_t)m * d * d + (int64_t)d * m; return 0;
      case SDB_DIFF_USER: *n = -1; return 0;
      case SDB_DIFF_USER_DIAG: *n = -1; return d == m ? 0 : 1;
    }
  }
  return 1;
}

static int system_create(const char* src, int fk, const double* fpar, int64_t nf, int gk,
                         const double* gpar, int64_t ng, int d, int m, sdb_system** out) {
  if (!out) return fail("sdb_system: out is NULL");
  if (d < 1 || m < 1 || d > 64 || m > 64) return fail("sdb_system: need 1 <= d, m <= 64");
  int64_t ef = 0, eg = 0;
  if (expected_params(fk, true, d, m, &ef)) return fail("sdb_system: bad drift kind / dimension");
  if (expected_params(gk, false, d, m, &eg)) return fail("sdb_system: bad diffusion kind / dimension");
  if (ef >= 0 && nf != ef) return fail("sdb_system: wrong number of drift parameters");
  if (eg >= 0 && ng != eg) return fail("sdb_system: wrong number of diffusion parameters");
  if ((fk == SDB_DRIFT_USER || gk == SDB_DIFF_USER || gk == SDB_DIFF_USER_DIAG) && (!src || !*src))
    return fail("sdb_system: user operator kinds need CUDA source");
  if (nf < 0 || ng < 0 || (nf > 0 && !fpar) || (ng > 0 && !gpar))
    return fail("sdb_system: parameter pointer is NULL");
  sdb_system* s = new sdb_system();
  s->fk = fk; s->gk = gk; s->d = d; s->m = m;
  s->fp.assign(fpar, fpar + nf);
  s->gp.assign(gpar, gpar + ng);
  if (s->fp.empty()) s->fp.push_back(0.0);
  if (s->gp.empty()) s->gp.push_back(0.0);
  if (src) s->src = src;
  s->inline_params =
      (s->fp.size() + s->gp.size()) * sizeof(double) + sizeof(SdbLoopArgs) + 64 <= kParamSpace;
  if (!s->inline_params) {
    cudaError_t e = cudaMalloc(&s->d_fp, s->fp.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_gp, s->gp.size() * sizeof(double));
    if (e == cudaSuccess)
      e = cudaMemcpy(s->d_fp, s->fp.data(), s->fp.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
      e = cudaMemcpy(s->d_gp, s->gp.data(), s->gp.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      delete s;
      return fail(std::string("sdb_system: device parameter upload: ") + cudaGetErrorString(e));
    }
  }
  *out = s;
  return 0;
}

extern "C" int sdb_system_builtin(int fk, const double* fpar, int64_t nf, int gk,
                                  const double* gpar, int64_t ng, int d, int m,
                                  sdb_system** out) {
  if (fk == SDB_DRIFT_USER || gk == SDB_DIFF_USER || gk == SDB_DIFF_USER_DIAG)
    return fail("sdb_system_builtin: user kinds need sdb_system_nvrtc");
  return system_create(nullptr, fk, fpar, nf, gk, gpar, ng, d, m, out);
}

extern "C" int sdb_system_nvrtc(const char* src, int fk, const double* fpar, int64_t nf, int gk,
                                const double* gpar, int64_t ng, int d, int m, sdb_system** out) {
  return system_create(src, fk, fpar, nf, gk, gpar, ng, d, m, out);
}

extern "C" int sdb_system_destroy(sdb_system* s) {
  if (!s) return 0;
  for (auto& kv : s->jit)
    if (kv.second.lib) cudaLibraryUnload(kv.second.lib);
  if (s->d_fp) cudaFree(s->d_fp);
  if (s->d_gp) cudaFree(s->d_gp);
  for (auto& kv : s->d_frags) cudaFree(kv.second);
  delete s;
  return 0;
}

static std::string type_of(const sdb_system* s, bool drift) {
  std::ostringstream o;
  const std::string ps_f =
      s->inline_params ? "PInline<" + std::to_string(s->fp.size()) + ">" : "PGlobal";
  const std::string ps_g =
      s->inline_params ? "PInline<" + std::to_string(s->gp.size()) + ">" : "PGlobal";
  if (drift) {
    switch (s->fk) {
      case SDB_DRIFT_LINEAR: o << "DriftLinear<" << s->d << ", " << ps_f << ">"; break;
      case SDB_DRIFT_KP4446: o << "DriftKP4446<" << ps_f << ">"; break;
      case SDB_DRIFT_KPS445: o << "DriftKPS445<" << ps_f << ">"; break;
      default: o << "DriftUser<" << s->d << ", " << ps_f << ">";
    }
  } else {
    switch (s->gk) {
      case SDB_DIFF_CONST: o << "DiffConst<" << s->d << ", " << s->m << ", " << ps_g << ">"; break;
      case SDB_DIFF_LINEAR_COLS:
        o << "DiffLinearCols<" << s->d << ", " << s->m << ", " << ps_g << ">"; break;
      case SDB_DIFF_KP4446: o << "DiffKP4446<" << ps_g << ">"; break;
      case SDB_DIFF_KPS445: o << "DiffKPS445<" << ps_g << ">"; break;
      case SDB_DIFF_DIAG_LINEAR: o << "DiffDiagLinear<" << s->d << ", " << ps_g << ">"; break;
      case SDB_DIFF_AFFINE_COLS:
        o << "DiffAffineCols<" << s->d << ", " << s->m << ", " << ps_g << ">"; break;
      case SDB_DIFF_USER_DIAG:
        o << "DiffUser<" << s->d << ", " << s->m << ", " << ps_g << ", true>"; break;
      default: o << "DiffUser<" << s->d << ", " << s->m << ", " << ps_g << ">";
    }
  }
  return o.str();
}

// Find (AOT) or build (NVRTC) the loop kernel for this system / scheme / noise source.
// ext: 0 = plain (ahead-of-time kernels allowed), 1 = resumable (SDB_RESUME), 2 = observing + resumable
static const char* ext_suffix(int ext) { return ext == 2 ? "|obs" : ext == 1 ? "|res" : ""; }
static const char* ext_defines(int ext) {
  return ext == 2 ? "#define SDB_OBSERVE 1\n#define SDB_RESUME 1\n" : ext == 1 ? "#define SDB_RESUME 1\n" : "";
}

static std::string aot_name(const void* fn) {
  const char* name = nullptr;
  if (cudaFuncGetName(&name, fn) != cudaSuccess || !name) { cudaGetLastError(); return "aot:?"; }
  return name;
}

static int loop_kernel(sdb_system* s, int scheme, int noise, const void** fn, int ext = 0,
                       std::string* label = nullptr) {
  std::string key = sdb_loop_key(scheme, s->fk, s->gk, s->d, s->m, noise, s->inline_params);
  const bool user_g = s->gk == SDB_DIFF_USER || s->gk == SDB_DIFF_USER_DIAG;
  const bool user = s->fk == SDB_DRIFT_USER || user_g;
  if (!user && !ext) {
    if (const void* f = aot_lookup(key)) {
      *fn = f;
      if (label) *label = aot_name(f);
      return 0;
    }
  }
  key += ext_suffix(ext);  // the observing / resumable variant of every kernel is compiled at run time
  std::lock_guard<std::mutex> lk(s->mu);
  auto it = s->jit.find(key);
  if (it == s->jit.end()) {
    // Small systems get their column loops unrolled (sdb_loop_body: SMALL).  When the user's code is
    // heavy that multiplies it: a kernel of 150 KB at 8 warps per SM is instruction-fetch bound (ncu:
    // no_instruction the top stall; +20 % from rolling the loops at d=7, m=5 with libm calls, -4 % for a
    // cheap d=m=3 system, -25 % for Euler).  So: build unrolled, and rebuild rolled only if the machine
    // code comes out larger than 100 KB.
    const char* lim = getenv("SDB_SMALL_LIMIT");
    JitKernel k;
    for (int attempt = 0; attempt < 2; ++attempt) {
      std::ostringstream src;
      src << ext_defines(ext);
      const char* mb = getenv("SDB_LOOP_MINBLOCKS");
      if (mb) src << "#define SDB_LOOP_MINBLOCKS " << atoi(mb) << "\n";
      if (lim) src << "#define SDB_SMALL_LIMIT " << atoi(lim) << "\n";
      else if (attempt == 1) {
        // rolled loops need fewer registers: cap them at 168 for 12 warps per SM instead of 8
        // (d=7, m=5 with libm calls: Ikpw +5 %, Iwik +36 %)
        src << "#define SDB_SMALL_LIMIT 0\n";
        if (!mb) src << "#define SDB_LOOP_MINBLOCKS 3\n";
      }
      if (s->fk == SDB_DRIFT_USER) src << "#define SDB_USER_DRIFT 1\n";
      if (user_g) src << "#define SDB_USER_DIFF 1\n";
      src << "#include \"sdb_kernels.cuh\"\n";
      if (user) src << "#line 1 \"user_source.cu\"\n" << s->src << "\n";
      src << "typedef " << type_of(s, true) << " F_jit;\n";
      src << "typedef " << type_of(s, false) << " G_jit;\n";
      src << "SDB_LOOP_KERNEL(sdb_jit_loop, " << scheme << ", F_jit, G_jit, " << noise << ")\n";
      size_t bytes = 0;
      JitKernel cand;
      if (jit_compile(src.str(), "sdb_jit_loop", &cand, &bytes)) return 1;
      if (attempt == 1) cudaLibraryUnload(k.lib);  // the unrolled build is dropped
      k = cand;
      // (Euler / Heun evaluate G once or twice per step and stay faster unrolled: 3.6e9 vs 3.1e9)
      if (lim || !user || s->d * s->m > 36 || scheme < SDB_SCHEME_SRK2 || bytes <= 100u * 1024u) break;
    }
    it = s->jit.emplace(key, k).first;
  }
  *fn = (const void*)it->second.kern;
  if (label) *label = "jit:" + key;
  return 0;
}

extern "C" const char* sdb_last_kernel(void) { return g_last_kernel.c_str(); }

// ---------------------------------------------------------------------------------------------
// sdb_integrate
// ---------------------------------------------------------------------------------------------
__global__ void sdb_k_status_init(int64_t* st) {
  st[0] = 0; st[1] = 0x7fffffffffffffffll; st[2] = 0; st[3] = 0x7fffffffffffffffll;
}

static int integrate_impl(sdb_system* s, const sdb_integrate_args* a, const SdbObsArgs* obs,
                          int64_t step0 = 0, double h_noise = 0.0);

extern "C" int sdb_integrate(sdb_system* s, const sdb_integrate_args* a) {
  return integrate_impl(s, a, nullptr);
}

extern "C" int sdb_integrate_observed(sdb_system* s, const sdb_integrate_args* a, const sdb_observer* h_obs,
                                      int n_obs, double* d_obs, int64_t stride_obs, int64_t stride_path) {
  const sdb_integrate_ext x{0, 0.0, n_obs, h_obs, d_obs, stride_obs, stride_path};
  return sdb_integrate_segment(s, a, &x);
}

static int make_observers(sdb_system* s, const sdb_observer* h_obs, int n_obs, double* d_obs,
                          int64_t stride_obs, int64_t stride_path, SdbObsArgs* o) {
  if (n_obs < 0 || n_obs > SDB_MAX_OBSERVERS)
    return fail("observers: need 0 <= n_obs <= SDB_MAX_OBSERVERS");
  if (!h_obs || !d_obs) return fail("observers: NULL observer buffer");
  static_assert(SDB_MAX_OBSERVERS == SDB_MAX_OBS, "sdb.h and sdb_kernels.cuh disagree");
  memset(o, 0, sizeof *o);
  o->n = n_obs;
  for (int i = 0; i < n_obs; ++i) {
    if (h_obs[i].kind < SDB_OBS_FIRST_ABOVE || h_obs[i].kind > SDB_OBS_LAGPROD)
      return fail("observers: unknown observer kind");
    if (h_obs[i].kind == SDB_OBS_LAGPROD &&
        !(h_obs[i].level >= 0.0 && h_obs[i].level <= (double)SDB_OBS_MAXLAG &&
          h_obs[i].level == (double)(int)h_obs[i].level))
      return fail("observers: the lag of a lagged-product observer must be an integer in [0, 32]");
    if (h_obs[i].comp < 0 || h_obs[i].comp >= s->d)
      return fail("observers: observer component out of range");
    o->kind[i] = h_obs[i].kind;
    o->comp[i] = h_obs[i].comp;
    o->level[i] = h_obs[i].level;
  }
  o->out = d_obs;
  o->sO = stride_obs;
  o->sP = stride_path;
  return 0;
}

extern "C" int sdb_integrate_segment(sdb_system* s, const sdb_integrate_args* a, const sdb_integrate_ext* x) {
  if (!s || !a || !x) return fail("sdb_integrate_segment: NULL argument");
  if (x->step0 < 0 || x->step0 + (int64_t)a->N - 1 > 0xffffffffll)
    return fail("sdb_integrate_segment: step counter out of range");
  if (x->h_noise < 0.0) return fail("sdb_integrate_segment: h_noise must be >= 0");
  if (x->step0 != 0 && a->scheme == SDB_SCHEME_KP2IS)
    return fail("sdb_integrate_segment: the two-step scheme KP2iS cannot be resumed");
  if (x->n_obs == 0) return integrate_impl(s, a, nullptr, x->step0, x->h_noise);
  SdbObsArgs o;
  if (make_observers(s, x->h_obs, x->n_obs, x->d_obs, x->obs_stride_obs, x->obs_stride_path, &o)) return 1;
  return integrate_impl(s, a, &o, x->step0, x->h_noise);
}

static int integrate_impl(sdb_system* s, const sdb_integrate_args* a, const SdbObsArgs* obs, int64_t step0,
                          double h_noise) {
  if (!s || !a) return fail("sdb_integrate: NULL argument");
  if (a->scheme < SDB_SCHEME_EULER || a->scheme > SDB_SCHEME_R2)
    return fail("sdb_integrate: unknown scheme");
  if (a->noise < SDB_NOISE_HOST || a->noise > SDB_NOISE_COMM)
    return fail("sdb_integrate: unknown noise source");
  if (a->N < 2) return fail("sdb_integrate: need at least two time points");
  if (a->paths < 1) return fail("sdb_integrate: need at least one path");
  if (a->save_every < 1) return fail("sdb_integrate: save_every must be >= 1");
  if (!a->h_tspan || !a->d_y0 || !a->d_y) return fail("sdb_integrate: NULL buffer");
  // additive noise never uses the repeated integrals (see NEED_IJ in sdb_kernels.cuh)
  const bool need_ij = (a->scheme == SDB_SCHEME_SRK2 || a->scheme == SDB_SCHEME_KP2IS) &&
                       s->gk != SDB_DIFF_CONST;
  // diagonal noise: I/J may be omitted, the kernel builds the diagonal from dW
  const bool diag = (s->gk == SDB_DIFF_DIAG_LINEAR || s->gk == SDB_DIFF_USER_DIAG) &&
                    a->scheme == SDB_SCHEME_SRK2;  // KP2iS does use the off-diagonal J
  if (a->noise == SDB_NOISE_HOST && (!a->d_dW || (need_ij && !a->d_IJ && !diag)))
    return fail("sdb_integrate: NOISE_HOST needs dW (and I/J for SRK2, KP2iS)");
  if ((a->noise == SDB_NOISE_KPW || a->noise == SDB_NOISE_WIK) && a->n_terms < 1)
    return fail("sdb_integrate: n_terms must be >= 1");
  if (a->scheme == SDB_SCHEME_KP2IS && !a->h_kp2is)
    return fail("sdb_integrate: KP2iS needs gam, al1, al2");
  cudaStream_t st = (cudaStream_t)a->stream;

  int noise = a->noise;
  int n_terms = a->n_terms;
  if (!need_ij && noise == SDB_NOISE_WIK) noise = SDB_NOISE_KPW;  // only dW is drawn
  // commutative noise: the Ikpw/Jkpw kernels with an empty series (A = 0, only dW is drawn)
  const bool comm = noise == SDB_NOISE_COMM;
  if (comm) { noise = SDB_NOISE_KPW; n_terms = 0; }
  const int ext = obs ? 2 : (step0 != 0 ? 1 : 0);  // which build of the kernel (see ext_defines)
  // The production path: linear system, SRI2/SRS2, on-device Philox + Ikpw/Jkpw (or Iwik/Jwik)
  // -> a fused kernel when one is instantiated for this (d, m, method).
  const SdbFusedEntry* fused = nullptr;
  SdbFusedEntry jit_entry{};
  int jit_rb = 4, jit_xr = 0;
  bool jit_comm = false;
  std::string label;  // what gets launched (sdb_last_kernel): an ahead-of-time kernel's name or "jit:<key>"
  const bool explicit_em = a->scheme == SDB_SCHEME_EULER || a->scheme == SDB_SCHEME_HEUN;
  // Commutative-noise integrals on a linear system: the dedicated kernel (sdb_fusedc.cuh) -- stage-2 sum
  // as W (W y) / 2 - (h/2) S y, no per-path products with I at all; (10,10) ahead of time, else NVRTC.
  if (comm && need_ij && a->scheme == SDB_SCHEME_SRK2 && s->fk == SDB_DRIFT_LINEAR &&
      s->gk == SDB_DIFF_LINEAR_COLS && !getenv("SDB_DISABLE_FUSED") && !getenv("SDB_DISABLE_FUSEDC") &&
      SdbFusedCDims(s->d, s->m).supported()) {
    if (!ext) {
      const char* vs = getenv("SDB_FUSED_VARIANT");
      const int want = vs ? atoi(vs) : 0;
      for (const auto& e : sdb_fused_registry())
        if (e.d == s->d && e.m == s->m && e.noise == SDB_NOISE_COMM && e.scheme == SDB_SCHEME_SRK2 &&
            (e.variant == want || (!fused && e.variant == 0)))
          fused = &e;
    }
    if (!fused && !getenv("SDB_DISABLE_FUSED_JIT") && s->d * s->m >= 16) {
      const SdbFusedCDims cd(s->d, s->m);
      const std::string key = std::string("fusedcjit") + ext_suffix(ext);
      std::lock_guard<std::mutex> lk(s->mu);
      auto it = s->jit.find(key);
      if (it == s->jit.end()) {
        std::ostringstream src;
        src << ext_defines(ext);
        src << "#include \"sdb_fusedc.cuh\"\nSDB_FUSEDC_KERNEL(sdb_jit_fused, " << s->d << ", " << s->m << ")\n";
        JitKernel k;
        if (jit_compile(src.str(), "sdb_jit_fused", &k)) return 1;
        it = s->jit.emplace(key, k).first;
      }
      jit_entry = SdbFusedEntry{s->d, s->m, 0, (const void*)it->second.kern, cd.smem_bytes(),
                                cd.table_doubles(), nullptr, cd.threads(), cd.threads(),
                                SDB_NOISE_COMM, SDB_SCHEME_SRK2};
      jit_comm = true;
      fused = &jit_entry;
      label = "jit:" + key;
    }
  }
  if (!fused && (a->scheme == SDB_SCHEME_SRK2 || explicit_em) && noise != SDB_NOISE_HOST &&
      s->fk == SDB_DRIFT_LINEAR && s->gk == SDB_DIFF_LINEAR_COLS &&
      n_terms <= SDB_FUSED_MAX_TERMS && !getenv("SDB_DISABLE_FUSED") && !ext) {
    const char* vs = getenv("SDB_FUSED_VARIANT");
    const int want = vs ? atoi(vs) : 0;
    for (const auto& e : sdb_fused_registry())
      if (e.d == s->d && e.m == s->m && e.noise == noise && e.scheme == a->scheme &&
          (e.variant == want || (!fused && e.variant == 0)))
        fused = &e;
  }
  // ANY drift (built-in or NVRTC source) with linear diffusion g_k = B_k y, SRI2/SRS2 + Ikpw/Jkpw on device
  // noise: the fused kernel with the drift evaluated per lane (SDB_FUSED_A2 -1 in sdb_fused.cuh) -- the
  // products with the B_k stay on the tensor instructions, only the two drift evaluations are user code.
  bool jit_ud = false, jit_aff = false;
  // affine diffusion g_k = B_k y + c_k (SDB_FUSED_AFFINE): the 8-warp kernel only, linear or any drift
  const bool aff_g = s->gk == SDB_DIFF_AFFINE_COLS;
  const bool ud_ok = !fused && a->scheme == SDB_SCHEME_SRK2 && noise != SDB_NOISE_HOST && need_ij &&
                     (s->fk != SDB_DRIFT_LINEAR || aff_g) && (s->gk == SDB_DIFF_LINEAR_COLS || aff_g) &&
                     n_terms <= SDB_FUSED_MAX_TERMS && !getenv("SDB_DISABLE_FUSED") &&
                     !getenv("SDB_DISABLE_FUSED_JIT") && s->d * s->m >= 16;
  if (ud_ok && noise == SDB_NOISE_KPW &&
      SdbFusedDims(s->d, s->m, 4, s->fk == SDB_DRIFT_LINEAR ? SDB_FUSED_A2 : -1).supported()) {
    // (a linear drift with affine diffusion keeps its table rows: A, A^2)
    const bool lin_f = s->fk == SDB_DRIFT_LINEAR;
    const SdbFusedDims dims(s->d, s->m, 4, lin_f ? SDB_FUSED_A2 : -1);
    const std::string key = std::string(lin_f ? "fusedjit" : "fusedjit|ud") + (aff_g ? "|aff" : "") + ext_suffix(ext);
    std::lock_guard<std::mutex> lk(s->mu);
    auto it = s->jit.find(key);
    if (it == s->jit.end()) {
      std::ostringstream src;
      src << ext_defines(ext);
      if (s->fk == SDB_DRIFT_USER) src << "#define SDB_USER_DRIFT 1\n";
      src << "#include \"sdb_kernels.cuh\"\n";
      if (s->fk == SDB_DRIFT_USER) src << "#line 1 \"user_source.cu\"\n" << s->src << "\n";
      src << "typedef " << type_of(s, true) << " F_jit;\n";
      if (aff_g) src << "#define SDB_FUSED_AFFINE 1\n";
      if (lin_f)
        src << "#include \"sdb_fused.cuh\"\nSDB_FUSED_KERNEL(sdb_jit_fused, " << s->d << ", " << s->m << ")\n";
      else
        src << "#define SDB_FUSED_A2 (-1)\n#define SDB_FUSED_DRIFT_T F_jit\n#include \"sdb_fused.cuh\"\n"
            << "SDB_FUSED_KERNEL_UD(sdb_jit_fused, " << s->d << ", " << s->m << ")\n";
      JitKernel k;
      if (jit_compile(src.str(), "sdb_jit_fused", &k)) return 1;
      it = s->jit.emplace(key, k).first;
    }
    jit_xr = lin_f ? SDB_FUSED_A2 : -1;
    jit_ud = !lin_f;
    jit_aff = aff_g;
    jit_entry = SdbFusedEntry{s->d, s->m, 0, (const void*)it->second.kern, dims.smem_bytes(),
                              dims.table_doubles() + (aff_g ? (size_t)s->d * s->m : 0), nullptr,
                              SdbFusedDims::THREADS, SdbFusedDims::THREADS, SDB_NOISE_KPW, SDB_SCHEME_SRK2};
    fused = &jit_entry;
    label = "jit:" + key;
  }
  // ... Iwik/Jwik and the larger shapes through the TMEM-resident kernel (sdb_fusedx.cuh, one lane per path)
  if (ud_ok && !fused && !aff_g && SdbFusedXDims(s->d, s->m, -1).supported()) {
    const SdbFusedXDims xdims(s->d, s->m, -1);
    const SdbFusedX2Dims x2dims(s->d, s->m, -1);   // two lanes per path where one warp per sub-partition is left
    const bool use_x2 = x2dims.supported() && !getenv("SDB_DISABLE_FUSEDX2");
    const bool wik = noise == SDB_NOISE_WIK;
    const std::string key = std::string(wik ? "fusedxjit|wik" : "fusedxjit|kpw") + (use_x2 ? "|x2" : "") + "|ud" +
                            ext_suffix(ext);
    std::lock_guard<std::mutex> lk(s->mu);
    auto it = s->jit.find(key);
    if (it == s->jit.end()) {
      std::ostringstream src;
      src << ext_defines(ext);
      if (s->fk == SDB_DRIFT_USER) src << "#define SDB_USER_DRIFT 1\n";
      src << "#include \"sdb_kernels.cuh\"\n";
      if (s->fk == SDB_DRIFT_USER) src << "#line 1 \"user_source.cu\"\n" << s->src << "\n";
      src << "typedef " << type_of(s, true) << " F_jit;\n";
      src << "#define SDB_FUSED_RB 2\n#define SDB_FUSED_VOL_MMA 0\n#define SDB_FUSED_A2 (-1)\n"
          << "#define SDB_FUSED_DRIFT_T F_jit\n#include \"" << (use_x2 ? "sdb_fusedx2.cuh" : "sdb_fusedx.cuh")
          << "\"\n" << (use_x2 ? "SDB_FUSEDX2_KERNEL_UD" : "SDB_FUSEDX_KERNEL_UD") << "(sdb_jit_fused, " << s->d
          << ", " << s->m << ", " << (wik ? 1 : 0) << ")\n";
      JitKernel k;
      if (jit_compile(src.str(), "sdb_jit_fused", &k)) return 1;
      it = s->jit.emplace(key, k).first;
    }
    jit_rb = 2;
    jit_xr = -1;
    jit_ud = true;
    jit_entry = SdbFusedEntry{s->d, s->m, 0, (const void*)it->second.kern,
                              use_x2 ? x2dims.smem_bytes() : xdims.smem_bytes(), xdims.sh.table_doubles(), nullptr,
                              use_x2 ? x2dims.threads() : xdims.threads(),
                              use_x2 ? x2dims.paths_per_block() : xdims.threads(), noise, SDB_SCHEME_SRK2};
    fused = &jit_entry;
    label = "jit:" + key;
  }
  // ... and the fused Euler / Heun kernels likewise (SDB_EH_XR -1 in sdb_fusedeh.cuh)
  if (!fused && explicit_em && noise != SDB_NOISE_HOST && s->fk != SDB_DRIFT_LINEAR &&
      s->gk == SDB_DIFF_LINEAR_COLS && !getenv("SDB_DISABLE_FUSED") && !getenv("SDB_DISABLE_FUSED_JIT") &&
      s->d * s->m >= (a->scheme == SDB_SCHEME_HEUN ? 48 : 96) &&
      SdbFusedEhDims(s->d, s->m, a->scheme == SDB_SCHEME_HEUN, -1).supported()) {
    const bool heun = a->scheme == SDB_SCHEME_HEUN;
    const SdbFusedEhDims edims(s->d, s->m, heun, -1);
    const std::string key = std::string(heun ? "fusedehjit|heun|ud" : "fusedehjit|euler|ud") + ext_suffix(ext);
    std::lock_guard<std::mutex> lk(s->mu);
    auto it = s->jit.find(key);
    if (it == s->jit.end()) {
      std::ostringstream src;
      src << ext_defines(ext);
      if (s->fk == SDB_DRIFT_USER) src << "#define SDB_USER_DRIFT 1\n";
      src << "#include \"sdb_kernels.cuh\"\n";
      if (s->fk == SDB_DRIFT_USER) src << "#line 1 \"user_source.cu\"\n" << s->src << "\n";
      src << "typedef " << type_of(s, true) << " F_jit;\n";
      src << "#define SDB_EH_XR (-1)\n#define SDB_FUSED_DRIFT_T F_jit\n#include \"sdb_fusedeh.cuh\"\n"
          << "SDB_FUSEDEH_KERNEL_UD(sdb_jit_fused, " << s->d << ", " << s->m << ", " << (heun ? 1 : 0) << ")\n";
      JitKernel k;
      if (jit_compile(src.str(), "sdb_jit_fused", &k)) return 1;
      it = s->jit.emplace(key, k).first;
    }
    jit_xr = -1;
    jit_ud = true;
    jit_entry = SdbFusedEntry{s->d, s->m, 0, (const void*)it->second.kern, edims.smem_bytes(),
                              edims.sh.table_doubles(), nullptr, edims.threads(), edims.threads(),
                              SDB_NOISE_KPW, a->scheme};
    fused = &jit_entry;
    label = "jit:" + key;
  }
  // No ahead-of-time instantiation for this (d, m, method): compile a fused kernel with NVRTC when
  // the shape is one it supports -- sdb_fused.cuh (8 warps per SM; Ikpw/Jkpw, m <= 10, d < 16) or
  // sdb_fusedx.cuh (Levy areas in tensor memory; Iwik/Jwik and larger shapes) -- else the generic loop.
  if (!fused && explicit_em && noise != SDB_NOISE_HOST && s->fk == SDB_DRIFT_LINEAR &&
      s->gk == SDB_DIFF_LINEAR_COLS && !getenv("SDB_DISABLE_FUSED") &&
      !getenv("SDB_DISABLE_FUSED_JIT") &&
      // measured crossover against the high-occupancy loop kernel (tools/bench_crossover.py):
      // Heun wins from d m ~ 48 (2.4x at (8,6)), Euler only for the largest shapes (1.5x at (10,10))
      s->d * s->m >= (a->scheme == SDB_SCHEME_HEUN ? 48 : 96) &&
      SdbFusedEhDims(s->d, s->m, a->scheme == SDB_SCHEME_HEUN).supported()) {
    const bool heun = a->scheme == SDB_SCHEME_HEUN;
    const SdbFusedEhDims edims(s->d, s->m, heun);
    const std::string key = std::string(heun ? "fusedehjit|heun" : "fusedehjit|euler") + ext_suffix(ext);
    std::lock_guard<std::mutex> lk(s->mu);
    auto it = s->jit.find(key);
    if (it == s->jit.end()) {
      std::ostringstream src;
      src << ext_defines(ext);
      src << "#include \"sdb_fusedeh.cuh\"\nSDB_FUSEDEH_KERNEL(sdb_jit_fused, " << s->d << ", " << s->m
          << ", " << (heun ? 1 : 0) << ")\n";
      JitKernel k;
      if (jit_compile(src.str(), "sdb_jit_fused", &k)) return 1;
      it = s->jit.emplace(key, k).first;
    }
    jit_entry = SdbFusedEntry{s->d, s->m, 0, (const void*)it->second.kern, edims.smem_bytes(),
                              edims.sh.table_doubles(), nullptr, edims.threads(), edims.threads(),
                              SDB_NOISE_KPW, a->scheme};
    fused = &jit_entry;
    label = "jit:" + key;
  }
  if (!fused && a->scheme == SDB_SCHEME_SRK2 && noise != SDB_NOISE_HOST &&
      s->fk == SDB_DRIFT_LINEAR && s->gk == SDB_DIFF_LINEAR_COLS &&
      n_terms <= SDB_FUSED_MAX_TERMS && !getenv("SDB_DISABLE_FUSED") &&
      !getenv("SDB_DISABLE_FUSED_JIT") && s->d * s->m >= 16) {  // tiny systems are generator-bound:
    const SdbFusedDims dims(s->d, s->m, 4, SDB_FUSED_A2);       // the high-occupancy loop kernel wins
    const SdbFusedXDims xdims(s->d, s->m);
    const bool use_v0 = noise == SDB_NOISE_KPW && dims.supported();
    const bool use_x = !use_v0 && xdims.supported();
    // one warp per sub-partition in sdb_fusedx.cuh (m >= 16): the two-lanes-per-path kernel instead
    const SdbFusedX2Dims x2dims(s->d, s->m);
    const bool use_x2 = use_x && x2dims.supported() && !getenv("SDB_DISABLE_FUSEDX2");
    if (use_v0 || use_x) {
      const std::string key =
          std::string(use_v0 ? "fusedjit" : (noise == SDB_NOISE_WIK ? "fusedxjit|wik" : "fusedxjit|kpw")) +
          (use_x2 ? "|x2" : "") + ext_suffix(ext);
      std::lock_guard<std::mutex> lk(s->mu);
      auto it = s->jit.find(key);
      if (it == s->jit.end()) {
        std::ostringstream src;
        src << ext_defines(ext);
        if (use_v0) {
          src << "#include \"sdb_fused.cuh\"\nSDB_FUSED_KERNEL(sdb_jit_fused, " << s->d << ", " << s->m << ")\n";
        } else {
          src << "#define SDB_FUSED_RB 2\n#define SDB_FUSED_VOL_MMA 0\n#include \""
              << (use_x2 ? "sdb_fusedx2.cuh" : "sdb_fusedx.cuh") << "\"\n"
              << (use_x2 ? "SDB_FUSEDX2_KERNEL" : "SDB_FUSEDX_KERNEL") << "(sdb_jit_fused, " << s->d << ", "
              << s->m << ", " << (noise == SDB_NOISE_WIK ? 1 : 0) << ")\n";
        }
        JitKernel k;
        if (jit_compile(src.str(), "sdb_jit_fused", &k)) return 1;
        it = s->jit.emplace(key, k).first;
      }
      if (use_v0) {
        jit_xr = SDB_FUSED_A2;
        jit_entry = SdbFusedEntry{s->d, s->m, 0, (const void*)it->second.kern, dims.smem_bytes(),
                                  dims.table_doubles(), nullptr, SdbFusedDims::THREADS,
                                  SdbFusedDims::THREADS, SDB_NOISE_KPW, SDB_SCHEME_SRK2};
      } else {
        jit_rb = 2;
        jit_xr = SDB_FUSED_A2;
        jit_entry = SdbFusedEntry{s->d, s->m, 0, (const void*)it->second.kern,
                                  use_x2 ? x2dims.smem_bytes() : xdims.smem_bytes(),
                                  xdims.sh.table_doubles(), nullptr,
                                  use_x2 ? x2dims.threads() : xdims.threads(),
                                  use_x2 ? x2dims.paths_per_block() : xdims.threads(), noise, SDB_SCHEME_SRK2};
      }
      fused = &jit_entry;
      label = "jit:" + key;
    }
  }
  const void* fn = nullptr;
  double* d_frag = nullptr;
  if (fused) {
    fn = fused->fn;
    std::lock_guard<std::mutex> lk(s->mu);
    auto it = s->d_frags.find(fn);
    if (it == s->d_frags.end()) {  // DMMA fragment tables of (A, B_k), built once per system and kernel
      SDB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused->smem));
      std::vector<double> tab(fused->table_doubles);
      if (fused->build) fused->build(s->fp.data(), s->gp.data(), tab.data());
      else if (jit_comm) SdbFusedCDims(s->d, s->m).build_tables(s->fp.data(), s->gp.data(), tab.data());
      else SdbFusedDims(s->d, s->m, jit_rb, jit_xr).build_tables(s->fp.data(), s->gp.data(), tab.data());
      if (jit_aff) {  // the constants c_k after the tables, as [j][row] (sdb_fused.cuh, SDB_FUSED_AFFINE)
        const size_t nb = (size_t)s->m * s->d * s->d;
        double* caff = tab.data() + tab.size() - (size_t)s->d * s->m;
        for (int j = 0; j < s->m; ++j)
          for (int r = 0; r < s->d; ++r) caff[(size_t)j * s->d + r] = s->gp[nb + (size_t)r * s->m + j];
      }
      double* d = nullptr;
      SDB_CUDA(cudaMalloc(&d, tab.size() * sizeof(double)));
      SDB_CUDA(cudaMemcpy(d, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
      it = s->d_frags.emplace(fn, d).first;
    }
    d_frag = it->second;
  } else if (loop_kernel(s, a->scheme, noise, &fn, ext, &label)) {
    return 1;
  }
  if (fused && label.empty()) label = aot_name(fn);
  g_last_kernel = label;
  if (getenv("SDB_TRACE"))
    fprintf(stderr, "sdb_integrate: scheme=%d d=%d m=%d noise=%d n_terms=%d paths=%lld steps=%d -> %s\n",
            a->scheme, s->d, s->m, a->noise, n_terms, (long long)a->paths, a->N - 1, label.c_str());

  // tspan, the per-step h = t[n+1]-t[n] with its square root and reciprocal root (uniform over
  // paths, so computed once here with IEEE sqrt / division) and the KP2iS coefficients go to
  // the device in one stream-ordered copy.
  const size_t nt = (size_t)a->N, ns = nt - 1;
  // KP2iS coefficients gam, al1, al2 (3 d); for a linear drift also (1 - diag(al2) h A)^-1 (d x d):
  // the implicit equation of the two-step scheme is then one fixed linear system (sdb_kernels.cuh)
  const bool kp_lin = a->scheme == SDB_SCHEME_KP2IS && s->fk == SDB_DRIFT_LINEAR;
  const size_t nd = (size_t)s->d;
  const size_t nk = a->scheme == SDB_SCHEME_KP2IS ? 3 * nd + (kp_lin ? nd * nd : 0) : 0;
  std::vector<double> aux(nt + 3 * ns + nk);
  for (size_t i = 0; i < nt; ++i) aux[i] = a->h_tspan[i];
  for (size_t i = 0; i < ns; ++i) {
    const double h = a->h_tspan[i + 1] - a->h_tspan[i];
    const double rh = sqrt(h);
    aux[nt + i] = h;
    aux[nt + ns + i] = rh;
    aux[nt + 2 * ns + i] = 1.0 / rh;
  }
  for (size_t i = 0; i < (nk ? 3 * nd : 0); ++i) aux[nt + 3 * ns + i] = a->h_kp2is[i];
  if (kp_lin) {
    const double hg = (a->h_tspan[a->N - 1] - a->h_tspan[0]) / (double)(a->N - 1);
    const int d = s->d;
    std::vector<double> Mx((size_t)d * 2 * d);  // [1 - diag(al2) h A | 1] -> [1 | inverse]
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) {
        Mx[(size_t)i * 2 * d + j] = (i == j ? 1.0 : 0.0) - a->h_kp2is[2 * d + i] * hg * s->fp[(size_t)i * d + j];
        Mx[(size_t)i * 2 * d + d + j] = i == j ? 1.0 : 0.0;
      }
    for (int c = 0; c < d; ++c) {
      int piv = c;
      for (int r = c + 1; r < d; ++r)
        if (fabs(Mx[(size_t)r * 2 * d + c]) > fabs(Mx[(size_t)piv * 2 * d + c])) piv = r;
      if (Mx[(size_t)piv * 2 * d + c] == 0.0)
        return fail("sdb_integrate: the implicit equation of KP2iS is singular (1 - al2 h A)");
      if (piv != c)
        for (int q = 0; q < 2 * d; ++q) std::swap(Mx[(size_t)c * 2 * d + q], Mx[(size_t)piv * 2 * d + q]);
      const double inv = 1.0 / Mx[(size_t)c * 2 * d + c];
      for (int q = 0; q < 2 * d; ++q) Mx[(size_t)c * 2 * d + q] *= inv;
      for (int r = 0; r < d; ++r) {
        if (r == c) continue;
        const double fct = Mx[(size_t)r * 2 * d + c];
        if (fct != 0.0)
          for (int q = 0; q < 2 * d; ++q) Mx[(size_t)r * 2 * d + q] -= fct * Mx[(size_t)c * 2 * d + q];
      }
    }
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) aux[nt + 3 * ns + 3 * nd + (size_t)i * d + j] = Mx[(size_t)i * 2 * d + d + j];
  }
  double* d_aux = nullptr;
  keep_pool_memory();
  SDB_CUDA(cudaMallocAsync(&d_aux, aux.size() * sizeof(double), st));
  SDB_CUDA(cudaMemcpyAsync(d_aux, aux.data(), aux.size() * sizeof(double), cudaMemcpyHostToDevice, st));

  SdbLoopArgs k;
  memset(&k, 0, sizeof k);
  k.P = a->paths; k.path_id0 = a->path_id0; k.seed = a->seed;
  k.N = a->N; k.save_every = a->save_every; k.n_terms = n_terms; k.flags = a->flags;
  k.h_global = h_noise > 0.0 ? h_noise : (a->h_tspan[a->N - 1] - a->h_tspan[0]) / (double)(a->N - 1);

  k.rtol = a->rtol;
  k.tspan = d_aux;
  k.y0 = a->d_y0; k.y0_sI = a->y0_stride_comp; k.y0_sP = a->y0_stride_path;
  k.y = a->d_y; k.y_sS = a->y_stride_save; k.y_sI = a->y_stride_comp; k.y_sP = a->y_stride_path;
  k.dW = a->d_dW; k.dW_sN = a->dW_stride_step; k.dW_sJ = a->dW_stride_comp; k.dW_sP = a->dW_stride_path;
  k.IJ = a->d_IJ; k.IJ_sN = a->IJ_stride_step; k.IJ_sR = a->IJ_stride_row;
  k.IJ_sC = a->IJ_stride_col; k.IJ_sP = a->IJ_stride_path;
  k.kp = nk ? d_aux + nt + 3 * ns : nullptr;
  k.status = a->d_status;
  if (obs) k.obs = *obs;
  k.obs.step0 = (uint32_t)step0;
  if (a->d_status) {
    sdb_k_status_init<<<1, 1, 0, st>>>(a->d_status);
    g_launches.fetch_add(1);
  }

  PGlobal gf{s->d_fp}, gg{s->d_gp};
  void* args[3];
  args[0] = &k;
  args[1] = s->inline_params ? (void*)s->fp.data() : (void*)&gf;
  args[2] = s->inline_params ? (void*)s->gp.data() : (void*)&gg;
  if (fused) {
    if (jit_ud) {  // (a, parameters of the drift functor, frag)
      args[2] = (void*)&d_frag;
    } else if (fused->scheme == SDB_SCHEME_SRK2 && fused->noise != SDB_NOISE_COMM) {
      args[1] = (void*)s->fp.data();  // the matrix A
      args[2] = (void*)&d_frag;
    } else {  // fused Euler / Heun and the commutative-noise kernel: (a, frag)
      args[1] = (void*)&d_frag;
    }
  }
  const unsigned block = fused ? (unsigned)fused->threads : SDB_BLOCK;
  const int64_t per_block = fused ? fused->paths_per_block : SDB_BLOCK;
  const unsigned grid = blocks_for(a->paths, per_block);
  const int rc = launch(fn, dim3(grid), dim3(block), args, st, fused ? fused->smem : 0);
  const cudaError_t fe = cudaFreeAsync(d_aux, st);  // stream-ordered: after the kernel has read it
  if (rc) return 1;
  if (fe != cudaSuccess) return fail(std::string("cudaFreeAsync: ") + cudaGetErrorString(fe));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// deltaW, repeated integrals
// ---------------------------------------------------------------------------------------------
extern "C" int sdb_delta_w(int64_t steps, int m, double h, int64_t paths, int64_t path_id0,
                           uint64_t seed, double* d_out, int64_t sN, int64_t sJ, int64_t sP,
                           void* stream) {
  if (steps < 0 || paths < 1 || m < 1 || !d_out) return fail("sdb_delta_w: bad argument");
  if (!(h > 0.0)) return fail("sdb_delta_w: h must be positive");
  if (steps == 0) return 0;
  SdbDeltaWArgs a{steps, paths, path_id0, seed, m, sqrt(h), d_out, sN, sJ, sP};
  void* args[1] = {&a};
  const int64_t total = steps * paths;
  return launch((const void*)sdb_k_delta_w, dim3(blocks_for(total, 256)), dim3(256), args,
                (cudaStream_t)stream);
}

SDB_REPINT_KERNEL(sdb_k_repint_1, 1)
SDB_REPINT_KERNEL(sdb_k_repint_2, 2)
SDB_REPINT_KERNEL(sdb_k_repint_3, 3)
SDB_REPINT_KERNEL(sdb_k_repint_4, 4)
SDB_REPINT_KERNEL(sdb_k_repint_5, 5)
SDB_REPINT_KERNEL(sdb_k_repint_6, 6)
SDB_REPINT_KERNEL(sdb_k_repint_7, 7)
SDB_REPINT_KERNEL(sdb_k_repint_8, 8)
SDB_REPINT_KERNEL(sdb_k_repint_9, 9)
SDB_REPINT_KERNEL(sdb_k_repint_10, 10)
SDB_REPINT_KERNEL(sdb_k_repint_20, 20)

static std::map<int, JitKernel> g_repint_jit;
static std::map<int, JitKernel> g_repintx_jit;  // key: 4 m + method
static std::mutex g_repint_mu;

static int repint_kernel(int m, const void** fn) {
  switch (m) {
    case 1: *fn = (const void*)sdb_k_repint_1; return 0;
    case 2: *fn = (const void*)sdb_k_repint_2; return 0;
    case 3: *fn = (const void*)sdb_k_repint_3; return 0;
    case 4: *fn = (const void*)sdb_k_repint_4; return 0;
    case 5: *fn = (const void*)sdb_k_repint_5; return 0;
    case 6: *fn = (const void*)sdb_k_repint_6; return 0;
    case 7: *fn = (const void*)sdb_k_repint_7; return 0;
    case 8: *fn = (const void*)sdb_k_repint_8; return 0;
    case 9: *fn = (const void*)sdb_k_repint_9; return 0;
    case 10: *fn = (const void*)sdb_k_repint_10; return 0;
    case 20: *fn = (const void*)sdb_k_repint_20; return 0;
  }
  std::lock_guard<std::mutex> lk(g_repint_mu);
  auto it = g_repint_jit.find(m);
  if (it == g_repint_jit.end()) {
    std::ostringstream src;
    src << "#include \"sdb_kernels.cuh\"\nSDB_REPINT_KERNEL(sdb_jit_repint, " << m << ")\n";
    JitKernel k;
    if (jit_compile(src.str(), "sdb_jit_repint", &k)) return 1;
    it = g_repint_jit.emplace(m, k).first;
  }
  *fn = (const void*)it->second.kern;
  return 0;
}

extern "C" int sdb_repeated_integrals(const sdb_repint_args* a) {
  if (!a) return fail("sdb_repeated_integrals: NULL argument");
  if (a->method != SDB_NOISE_KPW && a->method != SDB_NOISE_WIK && a->method != SDB_NOISE_COMM)
    return fail("sdb_repeated_integrals: method must be KPW, WIK or COMM");
  const bool comm = a->method == SDB_NOISE_COMM;  // = KPW with an empty series: A = 0, nothing drawn
  const int method = comm ? SDB_NOISE_KPW : a->method;
  if (a->m < 1 || a->m > 64) return fail("sdb_repeated_integrals: need 1 <= m <= 64");
  if (!comm && a->n_terms < 1) return fail("sdb_repeated_integrals: n_terms must be >= 1");
  if (a->steps < 0 || a->paths < 1) return fail("sdb_repeated_integrals: bad sizes");
  if (!(a->h > 0.0)) return fail("sdb_repeated_integrals: h must be positive");
  if (!a->d_dW || !a->d_IJ) return fail("sdb_repeated_integrals: NULL buffer");
  if ((a->d_X == nullptr) != (a->d_Y == nullptr))
    return fail("sdb_repeated_integrals: X and Y must be given together");
  if (a->d_X && method == SDB_NOISE_WIK && a->m > 1 && !a->d_Gt)
    return fail("sdb_repeated_integrals: host normals for WIK also need Gt");
  if (a->steps == 0) return 0;
  // Philox normals and a large even m: the TMEM-resident kernel (ahead of time for m = 20, else NVRTC)
  const void* fn = nullptr;
  size_t smem = 0;
  unsigned block = 128;
  if (!comm && !a->d_X && a->m >= 12 && a->n_terms <= SDB_FUSED_MAX_TERMS &&
      SdbRepintXDims(a->m).supported() && !getenv("SDB_DISABLE_FUSED")) {
    for (const auto& e : sdb_repintx_registry())
      if (e.m == a->m && e.noise == a->method) { fn = e.fn; smem = e.smem; block = (unsigned)e.threads; }
    if (!fn) {
      const int key = a->m * 4 + a->method;
      std::lock_guard<std::mutex> lk(g_repint_mu);
      auto it = g_repintx_jit.find(key);
      if (it == g_repintx_jit.end()) {
        std::ostringstream src;
        src << "#define SDB_FUSED_RB 2\n#define SDB_FUSED_VOL_MMA 0\n#include \"sdb_repintx.cuh\"\n"
            << "SDB_REPINTX_KERNEL(sdb_jit_repintx, " << a->m << ", " << (a->method == SDB_NOISE_WIK ? 1 : 0)
            << ")\n";
        JitKernel jk;
        if (jit_compile(src.str(), "sdb_jit_repintx", &jk)) return 1;
        it = g_repintx_jit.emplace(key, jk).first;
      }
      fn = (const void*)it->second.kern;
      smem = SdbRepintXDims(a->m).smem_bytes();
      block = (unsigned)SdbRepintXDims(a->m).threads();
    }
    SDB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  } else if (repint_kernel(a->m, &fn)) {
    return 1;
  }
  SdbRepintArgs k;
  memset(&k, 0, sizeof k);
  k.method = method; k.stratonovich = a->stratonovich; k.n_terms = comm ? 0 : a->n_terms; k.m = a->m;
  k.steps = a->steps; k.paths = a->paths; k.path_id0 = a->path_id0; k.seed = a->seed; k.h = a->h;
  k.dW = a->d_dW; k.dW_sN = a->dW_stride_step; k.dW_sJ = a->dW_stride_comp; k.dW_sP = a->dW_stride_path;
  k.X = a->d_X; k.Y = a->d_Y;
  k.XY_sT = a->XY_stride_term; k.XY_sN = a->XY_stride_step; k.XY_sJ = a->XY_stride_comp;
  k.XY_sP = a->XY_stride_path;
  k.Gt = a->d_Gt; k.Gt_sN = a->Gt_stride_step; k.Gt_sR = a->Gt_stride_pair; k.Gt_sP = a->Gt_stride_path;
  k.A = a->d_A; k.A_sN = a->A_stride_step; k.A_sE = a->A_stride_elem; k.A_sP = a->A_stride_path;
  k.IJ = a->d_IJ; k.IJ_sN = a->IJ_stride_step; k.IJ_sR = a->IJ_stride_row;
  k.IJ_sC = a->IJ_stride_col; k.IJ_sP = a->IJ_stride_path;
  void* args[1] = {&k};
  const int64_t total = a->steps * a->paths;
  return launch(fn, dim3(blocks_for(total, block)), dim3(block), args,
                (cudaStream_t)a->stream, smem);
}

// ---------------------------------------------------------------------------------------------
// moments / histograms
// ---------------------------------------------------------------------------------------------
#define SDB_MOM_CHUNK 4096
#define SDB_MOM_BLOCK 256

// partial[row][chunk] = (sum, sumsq) of 4096 consecutive paths, fixed tree order.
__global__ void __launch_bounds__(SDB_MOM_BLOCK)
sdb_k_moments_partial(const double* __restrict__ y, int64_t paths, int64_t nchunks,
                      double2* __restrict__ partial, unsigned long long* __restrict__ hist,
                      int bins, double lo, double inv_w) {
  extern __shared__ unsigned int sh_hist[];
  __shared__ double sh_s[SDB_MOM_BLOCK / 32], sh_q[SDB_MOM_BLOCK / 32];
  const int64_t row = blockIdx.y;
  const int64_t chunk = blockIdx.x;
  const double* src = y + row * paths;
  if (hist)
    for (int b = threadIdx.x; b < bins; b += blockDim.x) sh_hist[b] = 0u;
  __syncthreads();
  double s = 0.0, q = 0.0;
  const int64_t base = chunk * SDB_MOM_CHUNK;
#pragma unroll 4
  for (int i = 0; i < SDB_MOM_CHUNK / SDB_MOM_BLOCK; ++i) {
    const int64_t p = base + (int64_t)i * SDB_MOM_BLOCK + threadIdx.x;
    if (p < paths) {
      const double v = src[p];
      s += v;
      q = fma(v, v, q);
      if (hist) {
        const double f = (v - lo) * inv_w;
        if (f >= 0.0 && f < (double)bins) atomicAdd(&sh_hist[(int)f], 1u);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, o);
    q += __shfl_down_sync(0xffffffffu, q, o);
  }
  if ((threadIdx.x & 31) == 0) { sh_s[threadIdx.x >> 5] = s; sh_q[threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
#pragma unroll
    for (int w = 0; w < SDB_MOM_BLOCK / 32; ++w) { ts += sh_s[w]; tq += sh_q[w]; }
    partial[row * nchunks + chunk] = make_double2(ts, tq);
  }
  if (hist)
    for (int b = threadIdx.x; b < bins; b += blockDim.x)
      if (sh_hist[b]) atomicAdd(&hist[row * bins + b], (unsigned long long)sh_hist[b]);
}

// sums[row] += sum over chunks in chunk order (one thread per row: deterministic).
__global__ void sdb_k_moments_final(const double2* __restrict__ partial, int64_t rows,
                                    int64_t nchunks, double* __restrict__ sums) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  double s = 0.0, q = 0.0;
  for (int64_t c = 0; c < nchunks; ++c) {
    const double2 v = partial[row * nchunks + c];
    s += v.x;
    q += v.y;
  }
  sums[2 * row] += s;
  sums[2 * row + 1] += q;
}

extern "C" int sdb_moments(const double* d_y, int64_t n_saved, int d, int64_t paths, double* d_sums,
                           unsigned long long* d_hist, int bins, double lo, double hi,
                           void* stream) {
  if (!d_y || !d_sums || n_saved < 1 || d < 1 || paths < 1) return fail("sdb_moments: bad argument");
  if (d_hist && (bins < 1 || bins > 8192 || !(hi > lo))) return fail("sdb_moments: bad histogram spec");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = n_saved * d;
  const int64_t nchunks = (paths + SDB_MOM_CHUNK - 1) / SDB_MOM_CHUNK;
  if (rows > 65535) return fail("sdb_moments: too many (saved point, component) rows");
  double2* partial = nullptr;
  keep_pool_memory();
  SDB_CUDA(cudaMallocAsync(&partial, rows * nchunks * sizeof(double2), st));
  const double inv_w = d_hist ? (double)bins / (hi - lo) : 0.0;
  const size_t shm = d_hist ? bins * sizeof(unsigned int) : 0;
  sdb_k_moments_partial<<<dim3((unsigned)nchunks, (unsigned)rows), SDB_MOM_BLOCK, shm, st>>>(
      d_y, paths, nchunks, partial, d_hist, bins, lo, inv_w);
  SDB_CUDA(cudaGetLastError());
  sdb_k_moments_final<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(partial, rows, nchunks, d_sums);
  SDB_CUDA(cudaGetLastError());
  g_launches.fetch_add(2);
  SDB_CUDA(cudaFreeAsync(partial, st));
  return 0;
}


// ---- (count, mean, M2) partials with a fixed-order Chan merge -------------------------
// E[y^2] - E[y]^2 cancels catastrophically when |mean| >> std; here every chunk of 4096 paths is
// centred on its OWN mean (two passes over data that is L1/L2 resident after the first), and chunks
// are merged in chunk order with Chan's update  M2 = M2_a + M2_b + delta^2 n_a n_b / (n_a + n_b).
__global__ void __launch_bounds__(SDB_MOM_BLOCK)
sdb_k_welford_partial(const double* __restrict__ y, int64_t paths, int64_t nchunks,
                      double2* __restrict__ partial) {
  __shared__ double sh[SDB_MOM_BLOCK / 32];
  __shared__ double sh_mean;
  const int64_t row = blockIdx.y, chunk = blockIdx.x;
  const double* src = y + row * paths;
  const int64_t base = chunk * SDB_MOM_CHUNK;
  const int64_t cnt = paths - base < SDB_MOM_CHUNK ? paths - base : SDB_MOM_CHUNK;
  double v[SDB_MOM_CHUNK / SDB_MOM_BLOCK];
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < SDB_MOM_CHUNK / SDB_MOM_BLOCK; ++i) {
    const int64_t p = base + (int64_t)i * SDB_MOM_BLOCK + threadIdx.x;
    v[i] = p < paths ? src[p] : 0.0;
    s += v[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0;
#pragma unroll
    for (int w = 0; w < SDB_MOM_BLOCK / 32; ++w) ts += sh[w];
    sh_mean = ts / (double)cnt;
  }
  __syncthreads();
  const double mean = sh_mean;
  double q = 0.0;
#pragma unroll
  for (int i = 0; i < SDB_MOM_CHUNK / SDB_MOM_BLOCK; ++i) {
    const int64_t p = base + (int64_t)i * SDB_MOM_BLOCK + threadIdx.x;
    const double dlt = v[i] - mean;
    if (p < paths) q = fma(dlt, dlt, q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_down_sync(0xffffffffu, q, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tq = 0.0;
#pragma unroll
    for (int w = 0; w < SDB_MOM_BLOCK / 32; ++w) tq += sh[w];
    partial[row * nchunks + chunk] = make_double2(mean, tq);
  }
}

// out[row] = (count, mean, M2): Chan merge of the chunk partials in chunk order, then of the value
// already in out[row] (count 0 = empty), one thread per row: deterministic.
__global__ void sdb_k_welford_final(const double2* __restrict__ partial, int64_t rows, int64_t paths,
                                    int64_t nchunks, double* __restrict__ out) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  double n = out[3 * row], mean = out[3 * row + 1], m2 = out[3 * row + 2];
  for (int64_t c = 0; c < nchunks; ++c) {
    const double2 v = partial[row * nchunks + c];
    const int64_t left = paths - c * SDB_MOM_CHUNK;
    const double nb = (double)(left < SDB_MOM_CHUNK ? left : SDB_MOM_CHUNK);
    const double tot = n + nb, dlt = v.x - mean;
    mean += dlt * (nb / tot);
    m2 += v.y + dlt * dlt * (n * nb / tot);
    n = tot;
  }
  out[3 * row] = n; out[3 * row + 1] = mean; out[3 * row + 2] = m2;
}

extern "C" int sdb_moments_welford(const double* d_y, int64_t n_saved, int d, int64_t paths,
                                   double* d_out, void* stream) {
  if (!d_y || !d_out || n_saved < 1 || d < 1 || paths < 1) return fail("sdb_moments_welford: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = n_saved * d;
  const int64_t nchunks = (paths + SDB_MOM_CHUNK - 1) / SDB_MOM_CHUNK;
  if (rows > 65535) return fail("sdb_moments_welford: too many (saved point, component) rows");
  double2* partial = nullptr;
  keep_pool_memory();
  SDB_CUDA(cudaMallocAsync(&partial, rows * nchunks * sizeof(double2), st));
  sdb_k_welford_partial<<<dim3((unsigned)nchunks, (unsigned)rows), SDB_MOM_BLOCK, 0, st>>>(
      d_y, paths, nchunks, partial);
  SDB_CUDA(cudaGetLastError());
  sdb_k_welford_final<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(partial, rows, paths, nchunks, d_out);
  SDB_CUDA(cudaGetLastError());
  g_launches.fetch_add(2);
  SDB_CUDA(cudaFreeAsync(partial, st));
  return 0;
}

// cross[save][chunk][i*d+j] = sum over the 4096 paths of the chunk of y_i y_j: the chunk is staged
// through shared memory 256 paths at a time, thread t < d*d owns the pair (t / d, t % d).
__global__ void __launch_bounds__(256)
sdb_k_cross_partial(const double* __restrict__ y, int d, int64_t paths, int64_t nchunks,
                    double* __restrict__ partial) {
  extern __shared__ double tile[];  // [d][256]
  const int64_t save = blockIdx.y, chunk = blockIdx.x;
  const double* src = y + save * d * paths;
  const int t = threadIdx.x, i = t / d, j = t % d;
  const bool owner = t < d * d;
  double acc = 0.0;
  for (int b = 0; b < SDB_MOM_CHUNK / 256; ++b) {
    const int64_t p = chunk * SDB_MOM_CHUNK + (int64_t)b * 256 + t;
    for (int c = 0; c < d; ++c) tile[c * 256 + t] = p < paths ? src[c * paths + p] : 0.0;
    __syncthreads();
    if (owner) {
      const double* a = tile + i * 256;
      const double* bb = tile + j * 256;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;  // fixed association: deterministic
      for (int q = 0; q < 256; q += 4) {
        s0 = fma(a[q], bb[q], s0);
        s1 = fma(a[q + 1], bb[q + 1], s1);
        s2 = fma(a[q + 2], bb[q + 2], s2);
        s3 = fma(a[q + 3], bb[q + 3], s3);
      }
      acc += (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
  }
  if (owner) partial[(save * nchunks + chunk) * d * d + t] = acc;
}

__global__ void sdb_k_cross_final(const double* __restrict__ partial, int64_t n_saved, int dd,
                                  int64_t nchunks, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_saved * dd) return;
  const int64_t save = e / dd, t = e % dd;
  double s = 0.0;
  for (int64_t c = 0; c < nchunks; ++c) s += partial[(save * nchunks + c) * dd + t];
  out[e] += s;
}

extern "C" int sdb_cross_moments(const double* d_y, int64_t n_saved, int d, int64_t paths,
                                 double* d_out, void* stream) {
  if (!d_y || !d_out || n_saved < 1 || paths < 1) return fail("sdb_cross_moments: bad argument");
  if (d < 1 || d > 16) return fail("sdb_cross_moments: need 1 <= d <= 16");
  if (n_saved > 65535) return fail("sdb_cross_moments: too many saved points");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nchunks = (paths + SDB_MOM_CHUNK - 1) / SDB_MOM_CHUNK;
  double* partial = nullptr;
  keep_pool_memory();
  SDB_CUDA(cudaMallocAsync(&partial, (size_t)n_saved * nchunks * d * d * sizeof(double), st));
  sdb_k_cross_partial<<<dim3((unsigned)nchunks, (unsigned)n_saved), 256, (size_t)d * 256 * sizeof(double), st>>>(
      d_y, d, paths, nchunks, partial);
  SDB_CUDA(cudaGetLastError());
  const int64_t total = n_saved * d * d;
  sdb_k_cross_final<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(partial, n_saved, d * d, nchunks, d_out);
  SDB_CUDA(cudaGetLastError());
  g_launches.fetch_add(2);
  SDB_CUDA(cudaFreeAsync(partial, st));
  return 0;
}

extern "C" int sdb_copy2d_d2h_async(void* h_dst, int64_t dpitch, const void* d_src, int64_t spitch,
                                    int64_t width, int64_t height, void* stream) {
  if (!h_dst || !d_src || width < 0 || height < 0 || dpitch < width || spitch < width)
    return fail("sdb_copy2d_d2h_async: bad argument");
  if (width == 0 || height == 0) return 0;
  SDB_CUDA(cudaMemcpy2DAsync(h_dst, (size_t)dpitch, d_src, (size_t)spitch, (size_t)width,
                             (size_t)height, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// FP64 pipe peak (roofline denominator)
// ---------------------------------------------------------------------------------------------
#define SDB_PROBE_CHAINS 16
__global__ void __launch_bounds__(512) sdb_k_fp64_probe(int iters, double seed, double* sink) {
  double acc[SDB_PROBE_CHAINS];
  const double a = 1.0 + 1e-9 * seed, b = 1e-12 * (threadIdx.x + 1);
#pragma unroll
  for (int c = 0; c < SDB_PROBE_CHAINS; ++c) acc[c] = seed + c;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int c = 0; c < SDB_PROBE_CHAINS; ++c) acc[c] = fma(acc[c], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < SDB_PROBE_CHAINS; ++c) s += acc[c];
  if (s == 123.456) sink[0] = s;  // never true; keeps the chain alive
}

// One block of 512 threads per SM, 16 independent DFMA chains per thread (the configuration that
// reached 64 DFMA/clk/SM in tools/micro/ubench.cu).  `iters` sets the length of one launch
// (8 x 16 DFMAs per iteration): short launches give the burst figure, long ones the sustained one.
extern "C" int sdb_fp64_peak(int iters, double* tflops_out, void* stream) {
  if (!tflops_out || iters < 1) return fail("sdb_fp64_peak: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  SDB_CUDA(cudaGetDevice(&dev));
  SDB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* sink = nullptr;
  SDB_CUDA(cudaMalloc(&sink, sizeof(double)));
  cudaEvent_t e0, e1;
  SDB_CUDA(cudaEventCreate(&e0));
  SDB_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {  // rep 0 warms up
    SDB_CUDA(cudaEventRecord(e0, st));
    sdb_k_fp64_probe<<<sms, 512, 0, st>>>(iters, 1.0, sink);
    SDB_CUDA(cudaEventRecord(e1, st));
    SDB_CUDA(cudaEventSynchronize(e1));
    g_launches.fetch_add(1);
    float ms = 0.f;
    SDB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8.0 * SDB_PROBE_CHAINS * (double)iters * 512.0 * sms;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *tflops_out = best;
  return 0;
}

extern "C" int sdb_random_block(uint64_t seed, uint64_t path, uint32_t step, uint32_t slot,
                                uint32_t out[4]) {
  uint32_t r[4];
  sdb_random_block(SdbPhiloxKey(seed), path, step, slot, r);
  for (int i = 0; i < 4; ++i) out[i] = r[i];
  return 0;
}

extern "C" int sdb_philox_block(uint64_t seed, uint64_t path, uint32_t step, uint32_t slot,
                                uint32_t out[4]) {
  uint32_t r[4];
  sdb_philox4x32_10((uint32_t)path, (uint32_t)(path >> 32), step, slot, SdbPhiloxKey(seed), r);
  for (int i = 0; i < 4; ++i) out[i] = r[i];
  return 0;
}
