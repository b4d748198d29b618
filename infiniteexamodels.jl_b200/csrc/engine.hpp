// engine.hpp — interface between the C ABI (api.cpp) and the CUDA engine (engine.cu).
#pragma once
#include <string>

#include "plan.hpp"

namespace iexa {

enum Callback { CB_OBJ = 0, CB_GRAD = 1, CB_CONS = 2, CB_JAC = 3, CB_HESS = 4, CB_JPROD = 5, CB_JTPROD = 6, CB_HPROD = 7, CB__N = 8 };
// kernel slots of the specialised path: the five callbacks, jprod!, and the two phases of each scatter product
// (phase 0: groups whose single-writer outputs are plain stores; phase 1: everything else, atomics — plan.hpp)
enum KernelSlot { KS_JPROD = 5, KS_JTPROD0 = 6, KS_JTPROD1 = 7, KS_HPROD0 = 8, KS_HPROD1 = 9, KS_EVAL3 = 10, KS__N = 11 };
// NVRTC module of a kernel slot: 0 = the five callbacks (compiled at finalize), 1 = products, 2 = eval3 (compiled on first use)
inline int ks_set(int ks) { return ks < KS_JPROD ? 0 : ks < KS_EVAL3 ? 1 : 2; }
enum { GPROG_ALL = 6 };

struct Engine {
  virtual ~Engine() {}
  // all return IEXA_* status; err receives a message on failure
  virtual int structure(int which /*0 jac, 1 hess*/, void *rows, void *cols, int idx_bytes,
                        int memspace, void *stream, std::string &err) = 0;
  virtual int obj(const double *x, double *f_host, int memspace, void *stream, std::string &err) = 0;
  virtual int obj_device(const double *x_dev, double *f_dev, void *stream, std::string &err) = 0;
  virtual int grad(const double *x, double *g, int memspace, void *stream, std::string &err) = 0;
  virtual int cons(const double *x, double *c, int memspace, void *stream, std::string &err) = 0;
  virtual int jac(const double *x, double *vals, int memspace, void *stream, std::string &err) = 0;
  virtual int hess(const double *x, const double *y, double sigma, double *vals, int memspace,
                   void *stream, std::string &err) = 0;
  // cons! + jac_coord! + hess_coord! at one (x, y) in ONE launch
  virtual int eval3(const double *x, const double *y, double sigma, double *c, double *jvals, double *hvals, int memspace,
                    void *stream, std::string &err) = 0;
  virtual int jprod(const double *x, const double *v, double *Jv, int memspace, void *stream,
                    std::string &err) = 0;
  virtual int jtprod(const double *x, const double *v, double *Jtv, int memspace, void *stream,
                     std::string &err) = 0;
  virtual int hprod(const double *x, const double *y, const double *v, double sigma, double *Hv,
                    int memspace, void *stream, std::string &err) = 0;
  virtual int set_par(int64_t off, int64_t n, const double *vals, void *stream, bool device_sync, std::string &err) = 0;
  virtual int jac_rowptr(void *rowptr, int idx_bytes, int memspace, void *stream, std::string &err) = 0; // CSR row pointers (JAC_ROW_SORTED)
  virtual void device_bytes(int64_t *out) const = 0; // [columns, columns of the unsharded model, theta, theta full, programs + tables, host-path staging]
  virtual int get_column(int32_t col, double *out_host, std::string &err) = 0; // device copy of Plan::columns[col] (fp)
  virtual int host_register(void *p, size_t bytes, std::string &err) = 0;
  virtual int host_unregister(void *p, std::string &err) = 0;
  virtual int launches(int cb) const = 0;
  virtual int n_specialised() const = 0;
  virtual const char *note() const = 0; // why specialisation was skipped, or ""
};

// engine.cu; returns nullptr and fills err when no CUDA device / kernel image is usable
Engine *make_cuda_engine(Plan &plan, int device, uint32_t flags, std::string &err);

} // namespace iexa
