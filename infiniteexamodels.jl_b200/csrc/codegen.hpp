// codegen.hpp — run-time specialisation of the per-callback kernels (NVRTC, sm_100a).
//
// The reference stack specialises its evaluation code per expression-tree TYPE through Julia's
// JIT (ExaModels builds a distinct node type per tree).  The engine's
// equivalent: the register programs of all generators of one callback are printed into ONE
// straight-line CUDA kernel (switch over the generator id of the block), compiled for sm_100a
// with NVRTC at iexa_finalize, and launched exactly like the AOT interpreter kernel (same work
// table, same descriptors).  Registers replace the interpreter's local-memory register file.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "exec.hpp"
#include "plan.hpp"

namespace iexa {

class Specialiser {
 public:
  Specialiser();
  ~Specialiser();
  // builds one kernel per callback that has work; false + err if NVRTC/driver are unavailable
  bool build(const Plan &plan, const std::vector<GenD> &gens_host, int device, std::string &err);
  bool has(int cb) const { return cb >= 0 && cb < 5 && fn_[cb] != nullptr; }
  bool launch(int cb, int nblocks, const GenD *gens, const WorkItem *work, const double *x,
              const double *theta, const double *y, double sigma, double *out, double *partials,
              cudaStream_t st, std::string &err);
  int n_kernels() const { return n_kernels_; }
  // source of the generated translation unit (tests / DESIGN.md excerpts)
  static std::string generate_source(const Plan &plan);

 private:
  void *module_ = nullptr; // CUmodule
  void *fn_[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  int n_kernels_ = 0;
};

} // namespace iexa
