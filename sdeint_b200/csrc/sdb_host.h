// sdb_host.h -- host-side internals shared by sdb_capi.cu and the AOT translation units.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

// One ahead-of-time compiled kernel.  `key` encodes what it was instantiated for:
//   loop|s<scheme>|f<drift kind>|g<diff kind>|d<D>|m<M>|n<noise>|p<i|g>
//   repint|m<M>
struct SdbAotEntry {
  std::string key;
  const void* fn;
  size_t smem;  // dynamic shared memory the kernel needs (0 for the generic loops)
};

std::vector<SdbAotEntry>& sdb_aot_registry();

struct SdbAotRegistrar {
  SdbAotRegistrar(const std::string& key, const void* fn, size_t smem = 0) {
    sdb_aot_registry().push_back(SdbAotEntry{key, fn, smem});
  }
};

std::string sdb_loop_key(int scheme, int fk, int gk, int d, int m, int noise, bool inline_params);

// The fused linear SRI2/SRS2 + Philox/KPW kernel (sdb_fused.cuh) for one (d, m).
struct SdbFusedEntry {
  int d, m;
  int variant;                 // tuning variant (env SDB_FUSED_VARIANT; 0 = default)
  const void* fn;
  size_t smem;                 // dynamic shared memory
  size_t table_doubles;        // size of the DMMA fragment table built from (A, B)
  void (*build)(const double* A, const double* B, double* out);
  int threads;                 // block size
  int paths_per_block;         // sample paths integrated by one block
  int noise;                   // SDB_NOISE_KPW or SDB_NOISE_WIK: which repeated integrals it builds
  int scheme;                  // SDB_SCHEME_SRK2 (args: a, A, frag) or EULER / HEUN (args: a, frag)
};
std::vector<SdbFusedEntry>& sdb_fused_registry();
struct SdbFusedRegistrar {
  explicit SdbFusedRegistrar(const SdbFusedEntry& e) { sdb_fused_registry().push_back(e); }
};

// The TMEM-resident standalone repeated-integral kernel (sdb_repintx.cuh) for one (m, method).
struct SdbRepintXEntry {
  int m, noise;
  const void* fn;
  size_t smem;
  int threads;
};
std::vector<SdbRepintXEntry>& sdb_repintx_registry();
struct SdbRepintXRegistrar {
  explicit SdbRepintXRegistrar(const SdbRepintXEntry& e) { sdb_repintx_registry().push_back(e); }
};
