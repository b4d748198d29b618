// codegen.hpp — run-time specialisation of the per-callback kernels (NVRTC, sm_100a).
//
// The reference stack specialises its evaluation code per expression-tree TYPE through Julia's
// JIT (ExaModels builds a distinct node type per tree).  The engine's
// equivalent: the fused register programs (plan.hpp: Group) of one callback are printed into ONE
// straight-line CUDA kernel (switch over the group id of the block), compiled for sm_100a with
// NVRTC at iexa_finalize and launched once per callback.  Registers replace the interpreter's
// local-memory register file; shared memory turns the AoS COO layout into coalesced stores.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "exec.hpp"
#include "plan.hpp"

namespace iexa {

// entries of the generated kernels' __constant__ table (values are resolved per rank at load)
enum { CI_K0 = 0, CI_K1, CI_IDX_BASE, CI_ICOL_PTR, CI_FCOL_PTR, CI_MEM_ROWLOC, CI_MEM_OUT, CI_CLS_NBLK, CI_CLS_ITAB, CI_CLS_DTAB };
struct CiEntry {
  int kind, group, a, b;
};

struct GeneratedSource {
  std::string text;
  std::vector<CiEntry> ci;
  std::vector<int> groups_of[5]; // groups with work, per callback
  size_t smem_bytes[5] = {0, 0, 0, 0, 0};
};

int spec_block(); // threads per block of the specialised kernels (env IEXA_BLOCK, default 128)
GeneratedSource generate_source(const Plan &plan);
bool compile_cubin(const std::string &src, std::vector<char> &cubin, std::string &err);

class Specialiser {
 public:
  Specialiser();
  ~Specialiser();
  // col_dev_ptr[c]: device address of Plan::columns[c] (nullptr for iota columns)
  // may switch the plan to class mode (shape canonicalisation) when the source would exceed the budget
  bool build(Plan &plan, const std::vector<const void *> &col_dev_ptr, std::string &err);
  bool has(int cb) const { return cb >= 0 && cb < 5 && fn_[cb] != nullptr; }
  const std::vector<int> &groups_of(int cb) const { return groups_of_[cb]; }
  bool launch(int cb, int nblocks, const WorkItem *work, const double *x, const double *theta,
              const double *y, double sigma, double *out, double *partials, cudaStream_t st,
              std::string &err);
  int n_kernels() const { return n_kernels_; }
  static std::string generate_source(const Plan &plan);

 private:
  void *module_ = nullptr; // CUmodule
  std::vector<unsigned long long> dev_allocs_; // instance tables of class groups
  void *fn_[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t smem_[5] = {0, 0, 0, 0, 0};
  std::vector<int> groups_of_[5];
  int n_kernels_ = 0;
};

} // namespace iexa
