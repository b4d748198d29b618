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

#include "engine.hpp"
#include "exec.hpp"
#include "plan.hpp"

namespace iexa {

// entries of the generated kernels' __constant__ table (values are resolved per rank at load)
enum { CI_K0 = 0, CI_K1, CI_IDX_BASE, CI_ICOL_PTR, CI_FCOL_PTR, CI_MEM_ROWLOC, CI_MEM_OUT, CI_CLS_NBLK, CI_CLS_ITAB, CI_CLS_DTAB,
       CI_SB,         // supports per block of (group, callback a)
       CI_SCHED };    // group = -1, a = callback, b = 0: NB (arithmetic grid rows), 1: #work-table items, 2 + j: group id of slot j
struct CiEntry {
  int kind, group, a, b;
};

// Block schedule of one callback.  "Big" groups get NB blocks each and blockIdx.x -> (group, block) is ARITHMETIC
// (slot = blockIdx.x % nbig, block = blockIdx.x / nbig): no dependent work-table load at the top of the block.
// Every big group is cut into the SAME number of blocks — sb supports per block, a multiple of 32 chosen per group —
// so consecutive blocks of different groups sit at the same RELATIVE position of their support ranges (groups that
// walk the same supports read the same part of x at about the same time: L2 reuse).  Small groups (less than a warp
// of supports per block at that cut) and shape-class groups stay on the work table and are dispatched after the
// closed-form blocks.  Measured on B200 (config 3): the table load showed up as ~10 % of the stall samples of every
// callback in ncu's source view, but removing it is time-neutral (0.5482 vs 0.5475 ms per eval): other resident
// blocks were already covering it.  IEXA_SCHED=t keeps everything on the table.
struct CbSchedule {
  std::vector<int> big, sb, small;
  int64_t NB = 0;      // blocks per big group (= grid rows walked arithmetically)
  int64_t ntable = 0;  // work-table items (small and shape-class groups)
  int64_t gx = 1, gy = 0; // launch grid: x = big-group slot (fastest in dispatch order), y = block row
  int64_t nblocks() const { return gx * gy; }
};
CbSchedule make_schedule(const Plan &plan, const std::vector<int> &groups);

// The kernels live in two NVRTC modules: the five NLPModels callbacks every solver uses (set 0, compiled at
// iexa_finalize) and the matrix-free products jprod! / jtprod! / hprod! (set 1, compiled on their first call: MadNLP and
// Ipopt never call them, so they cost those solvers no build time).  Arrays below are indexed by kernel slot (engine.hpp).
struct GeneratedSource {
  CbSchedule sched[KS__N];
  std::string text;
  std::vector<CiEntry> ci;
  std::vector<int> groups_of[KS__N]; // groups with work, per kernel slot
  size_t smem_bytes[KS__N] = {0};
};

int spec_block(); // threads per block of the specialised kernels (env IEXA_BLOCK, default 128)
int class_chunk(); // instances of a shape class per block (env IEXA_CLASS_CHUNK, default 8)
GeneratedSource generate_source(const Plan &plan, int set = 0);
bool compile_cubin(const std::string &src, std::vector<char> &cubin, std::string &err);
void cache_stats(int *nvrtc_compiles, int *disk_hits); // of this process

class Specialiser {
 public:
  Specialiser();
  ~Specialiser();
  // col_dev_ptr[c]: device address of Plan::columns[c] (nullptr for iota columns)
  // may switch the plan to class mode (shape canonicalisation) when the source would exceed the budget
  bool build(Plan &plan, const std::vector<const void *> &col_dev_ptr, std::string &err);
  // second module: jprod! / jtprod! / hprod! kernels of the SAME groups (never regroups the plan)
  bool build_products(Plan &plan, const std::vector<const void *> &col_dev_ptr, std::string &err);
  bool products_built() const { return module_[1] != nullptr; }
  // third module: the fused cons! + jac_coord! + hess_coord! kernel (iexa_eval3), compiled on its first use
  bool build_eval3(Plan &plan, const std::vector<const void *> &col_dev_ptr, std::string &err);
  bool has(int ks) const { return ks >= 0 && ks < KS__N && fn_[ks] != nullptr; }
  const std::vector<int> &groups_of(int ks) const { return groups_of_[ks]; }
  const CbSchedule &schedule(int ks) const { return sched_[ks]; }
  bool launch(int ks, const WorkItem *work, const double *x, const double *theta,
              const double *y, const double *v, double sigma, double *out, double *partials, cudaStream_t st,
              std::string &err, double *out2 = nullptr, double *out3 = nullptr);
  int n_kernels() const { return n_kernels_; }
  static std::string generate_source(const Plan &plan, int set = 0);

 private:
  bool load_set(Plan &plan, const std::vector<const void *> &col_dev_ptr, int set, bool allow_regroup, std::string &err);
  void *module_[3] = {nullptr, nullptr, nullptr}; // CUmodule per kernel set (engine.hpp: ks_set)
  std::vector<unsigned long long> dev_allocs_; // instance tables of class groups
  void *fn_[KS__N] = {nullptr};
  size_t smem_[KS__N] = {0};
  std::vector<int> groups_of_[KS__N];
  CbSchedule sched_[KS__N];
  int n_kernels_ = 0;
};

} // namespace iexa
