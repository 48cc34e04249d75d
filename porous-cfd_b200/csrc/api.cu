// C-ABI entry points that dispatch between the jet-GEMM engines, plus version / device queries.
#include "common.cuh"

extern "C" {
int pcfd_ffma_jet_linear_fwd(const float*, int64_t, int32_t, const pcfd_intrans_t*, const float*, int32_t, const float*,
                             const float*, int32_t, float*, int64_t, int32_t, int32_t, int64_t, int64_t, int32_t,
                             int32_t, void*);
int pcfd_ffma_jet_linear_bwd_dx(const float*, int64_t, int32_t, const float*, int32_t, const float*, int64_t, int32_t,
                                const pcfd_intrans_t*, float*, int64_t, int32_t, float*, int32_t, int32_t, int64_t,
                                int64_t, int32_t, int32_t, void*);
int pcfd_ffma_jet_linear_bwd_dw(const float*, int64_t, int32_t, const float*, int64_t, int32_t, const pcfd_intrans_t*,
                                float*, int32_t, float*, float*, int32_t, int32_t, int64_t, int64_t, int32_t, int32_t,
                                void*, size_t, void*);
size_t pcfd_ffma_dw_workspace_bytes(int32_t, int64_t, int64_t, int32_t, int32_t);
int pcfd_dw_finish(const float*, int, const float*, int32_t, float*, int32_t, float*, float*, int32_t, int64_t, int64_t,
                   int32_t, int32_t, float*, const float*, int, void*);
int pcfd_thin_fwd_kind(const pcfd_intrans_t*, int32_t, int64_t, int32_t, int32_t);
int pcfd_thin_jet_linear_fwd(const float*, int64_t, int32_t, const pcfd_intrans_t*, const float*, int32_t, const float*,
                             const float*, int32_t, float*, int64_t, int32_t, int32_t, int64_t, int64_t, int32_t, int32_t,
                             void*);
int pcfd_thin_dx_kind(const pcfd_intrans_t*, const float*, int32_t, int64_t, int32_t, int32_t);
int pcfd_thin_jet_linear_bwd_dx(const float*, int64_t, int32_t, const float*, int32_t, const float*, int64_t, int32_t,
                                const pcfd_intrans_t*, float*, int64_t, int32_t, float*, int32_t, int32_t, int64_t, int64_t,
                                int32_t, int32_t, void*);
int pcfd_small_rows_supported_dw(const pcfd_intrans_t*, int32_t, int64_t, int32_t, int32_t);
int pcfd_small_rows_bwd_dw(const float*, int32_t, const float*, int32_t, float*, int32_t, float*, float*, int32_t, int64_t,
                           int64_t, int32_t, int32_t, void*);
int pcfd_ws_supported_fwd(const float*, int64_t, int32_t, const float*, int32_t, const float*, int64_t, int32_t, int32_t,
                          int64_t, int32_t, int32_t);
int pcfd_ws_supported_dx(const float*, int64_t, int32_t, const float*, int32_t, const float*, int64_t, int32_t, const float*,
                         int64_t, int32_t, int32_t, int64_t, int32_t, int32_t);
int pcfd_ws_jet_linear_bwd_dx(const float*, int64_t, int32_t, const float*, int32_t, const float*, int64_t, int32_t,
                              const pcfd_intrans_t*, float*, int64_t, int32_t, float*, int32_t, int32_t, int64_t, int64_t,
                              int32_t, int32_t, void*);
int pcfd_ws_supported_dw(const float*, int64_t, int32_t, const float*, int64_t, int32_t, int32_t, int64_t, int32_t, int32_t);
size_t pcfd_ws_dw_workspace_bytes(int32_t, int64_t, int64_t, int32_t, int32_t);
int pcfd_ws_jet_linear_bwd_dw_partials(const float*, int64_t, int32_t, const float*, int64_t, int32_t,
                                       const pcfd_intrans_t*, int32_t, int64_t, int64_t, int32_t, int32_t, void*, int, int*,
                                       void*);
int pcfd_ws_jet_linear_fwd(const float*, int64_t, int32_t, const pcfd_intrans_t*, const float*, int32_t, const float*,
                           const float*, int32_t, float*, int64_t, int32_t, int32_t, int64_t, int64_t, int32_t,
                           int32_t, void*);
int pcfd_ws_fwd1_enabled(int32_t);
int pcfd_ws_jet_linear_fwd1(const float*, int32_t, const pcfd_intrans_t*, const float*, int32_t, const float*, const float*,
                            int32_t, float*, int32_t, int64_t, int64_t, int32_t, int32_t, void*);
}

// The library keeps no mutable state: which kernel family executes a layer is a pure function of the call's shapes and
// pointers (pcfd_jet_linear_engine), the device check is repeated per call (two cached runtime queries).
static int ensure_arch() { return pcfd::check_sm100(); }

extern "C" int pcfd_abi_version(void) { return PCFD_ABI_VERSION; }

extern "C" int pcfd_device_arch(int* cc_out_host) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return PCFD_ERR_ARCH;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return PCFD_ERR_ARCH;
  if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return PCFD_ERR_ARCH;
  if (cc_out_host) *cc_out_host = major * 10 + minor;
  return PCFD_OK;
}

static inline bool bad_jet(int cj, int64_t rows, int k, int n) { return !pcfd::valid_cj(cj) || rows <= 0 || k <= 0 || n <= 0; }

extern "C" int pcfd_jet_linear_fwd(const float* zin, int64_t zin_ps, int32_t ldzin, const pcfd_intrans_t* tin,
                                   const float* w, int32_t ldw, const float* bias, const float* cvec, int32_t ldcvec,
                                   float* zout, int64_t zout_ps, int32_t ldzout, int32_t cj, int64_t rows,
                                   int64_t rows_per_geom, int32_t k, int32_t n, void* stream) {
  if (!zin || !w || !zout || bad_jet(cj, rows, k, n) || ldzin < k || ldw < k || ldzout < n) return PCFD_ERR_ARG;
  if ((cvec || (tin && tin->escale)) && rows_per_geom <= 0) return PCFD_ERR_ARG;
  if (tin && tin->drop_p > 0.f && !tin->seed_dev) return PCFD_ERR_ARG;
  int rc = ensure_arch();
  if (rc) return rc;
  if (pcfd_ws_supported_fwd(zin, zin_ps, ldzin, w, ldw, zout, zout_ps, ldzout, cj, rows, k, n)) {
    if (cj == 1 && pcfd_ws_fwd1_enabled(k))   // value-only layer: A operand through tensor memory
      return pcfd_ws_jet_linear_fwd1(zin, ldzin, tin, w, ldw, bias, cvec, ldcvec, zout, ldzout, rows, rows_per_geom, k, n,
                                     stream);
    return pcfd_ws_jet_linear_fwd(zin, zin_ps, ldzin, tin, w, ldw, bias, cvec, ldcvec, zout, zout_ps, ldzout, cj, rows,
                                  rows_per_geom, k, n, stream);
  }
  // shapes a tensor-core tile cannot use: last layer (n <= 8), first layer (k <= 16), per-geometry rows (<= 32)
  if (pcfd_thin_fwd_kind(tin, cj, rows, k, n) >= 0)
    return pcfd_thin_jet_linear_fwd(zin, zin_ps, ldzin, tin, w, ldw, bias, cvec, ldcvec, zout, zout_ps, ldzout, cj, rows,
                                    rows_per_geom, k, n, stream);
  return pcfd_ffma_jet_linear_fwd(zin, zin_ps, ldzin, tin, w, ldw, bias, cvec, ldcvec, zout, zout_ps, ldzout, cj, rows,
                                  rows_per_geom, k, n, stream);
}

extern "C" int pcfd_jet_linear_bwd_dx(const float* gzout, int64_t gzout_ps, int32_t ldgzout, const float* w,
                                      int32_t ldw, const float* zin, int64_t zin_ps, int32_t ldzin,
                                      const pcfd_intrans_t* tin, float* gzin, int64_t gzin_ps, int32_t ldgzin,
                                      float* gescale, int32_t ldgescale, int32_t cj, int64_t rows,
                                      int64_t rows_per_geom, int32_t k, int32_t n, void* stream) {
  if (!gzout || !w || !zin || !gzin || bad_jet(cj, rows, k, n) || ldgzout < n || ldw < k || ldzin < k || ldgzin < k)
    return PCFD_ERR_ARG;
  if ((gescale || (tin && tin->escale)) && rows_per_geom <= 0) return PCFD_ERR_ARG;
  if (tin && tin->drop_p > 0.f && !tin->seed_dev) return PCFD_ERR_ARG;
  int rc = ensure_arch();
  if (rc) return rc;
  if (pcfd_thin_dx_kind(tin, gescale, cj, rows, k, n) >= 0)
    return pcfd_thin_jet_linear_bwd_dx(gzout, gzout_ps, ldgzout, w, ldw, zin, zin_ps, ldzin, tin, gzin, gzin_ps, ldgzin,
                                       gescale, ldgescale, cj, rows, rows_per_geom, k, n, stream);
  if (pcfd_ws_supported_dx(gzout, gzout_ps, ldgzout, w, ldw, zin, zin_ps, ldzin, gzin, gzin_ps, ldgzin, cj, rows, k, n))
    return pcfd_ws_jet_linear_bwd_dx(gzout, gzout_ps, ldgzout, w, ldw, zin, zin_ps, ldzin, tin, gzin, gzin_ps, ldgzin,
                                     gescale, ldgescale, cj, rows, rows_per_geom, k, n, stream);
  return pcfd_ffma_jet_linear_bwd_dx(gzout, gzout_ps, ldgzout, w, ldw, zin, zin_ps, ldzin, tin, gzin, gzin_ps, ldgzin,
                                     gescale, ldgescale, cj, rows, rows_per_geom, k, n, stream);
}

extern "C" int pcfd_jet_linear_bwd_dw(const float* gzout, int64_t gzout_ps, int32_t ldgzout, const float* zin,
                                      int64_t zin_ps, int32_t ldzin, const pcfd_intrans_t* tin, float* gw, int32_t ldgw,
                                      float* gbias, float* gcvec, int32_t ldgcvec, int32_t cj, int64_t rows,
                                      int64_t rows_per_geom, int32_t k, int32_t n, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  if (!gzout || !zin || !workspace || bad_jet(cj, rows, k, n) || ldgzout < n || ldzin < k || (gw && ldgw < k))
    return PCFD_ERR_ARG;
  if (tin && tin->escale && rows_per_geom <= 0) return PCFD_ERR_ARG;
  if (tin && tin->drop_p > 0.f && !tin->seed_dev) return PCFD_ERR_ARG;
  int rc = ensure_arch();
  if (rc) return rc;
  if (workspace_bytes < pcfd_jet_linear_bwd_dw_workspace_bytes(cj, rows, rows_per_geom, k, n)) return PCFD_ERR_WORKSPACE;
  if (pcfd_small_rows_supported_dw(tin, cj, rows, k, n))
    return pcfd_small_rows_bwd_dw(gzout, ldgzout, zin, ldzin, gw, ldgw, gbias, gcvec, ldgcvec, rows, rows_per_geom, k, n,
                                  stream);
  if (gw != nullptr && pcfd_ws_supported_dw(gzout, gzout_ps, ldgzout, zin, zin_ps, ldzin, cj, rows, k, n)) {
    int splits = 0;
    const int fused_sums = gbias != nullptr && gcvec == nullptr;     // the dW kernel sums plane 0 of gzout on the way
    rc = pcfd_ws_jet_linear_bwd_dw_partials(gzout, gzout_ps, ldgzout, zin, zin_ps, ldzin, tin, cj, rows, rows_per_geom, k,
                                            n, workspace, fused_sums, &splits, stream);
    if (rc) return rc;
    float* partial = reinterpret_cast<float*>(workspace);
    float* colsum = partial + (size_t)splits * n * k;
    float* tmp = colsum + (size_t)splits * n;
    return pcfd_dw_finish(partial, splits, gzout, ldgzout, gw, ldgw, gbias, gcvec, ldgcvec, rows, rows_per_geom, k, n, tmp,
                          fused_sums ? colsum : nullptr, splits, stream);
  }
  return pcfd_ffma_jet_linear_bwd_dw(gzout, gzout_ps, ldgzout, zin, zin_ps, ldzin, tin, gw, ldgw, gbias, gcvec, ldgcvec,
                                     cj, rows, rows_per_geom, k, n, workspace, workspace_bytes, stream);
}

extern "C" size_t pcfd_jet_linear_bwd_dw_workspace_bytes(int32_t cj, int64_t rows, int64_t rows_per_geom, int32_t k,
                                                         int32_t n) {
  if (!pcfd::valid_cj(cj) || rows <= 0 || k <= 0 || n <= 0) return 0;
  size_t need = pcfd_ffma_dw_workspace_bytes(cj, rows, rows_per_geom, k, n);
  const size_t u = pcfd_ws_dw_workspace_bytes(cj, rows, rows_per_geom, k, n);
  if (u > need) need = u;
  return need;
}

// Which kernel family pcfd_jet_linear_{fwd,bwd_dx,bwd_dw} run for a call with these arguments (pass: 0 forward, 1 dX,
// 2 dW): a pure query, nothing is launched.  A caller that expects its wide layers on the tensor cores asserts on it
// (bench.py, tests) instead of discovering a 5x slower fallback in a profile.
extern "C" int pcfd_jet_linear_engine(int32_t pass, const float* a, int64_t a_ps, int32_t lda, const float* w, int32_t ldw,
                                      const float* b, int64_t b_ps, int32_t ldb, const pcfd_intrans_t* tin,
                                      int32_t has_gescale, int32_t cj, int64_t rows, int32_t k, int32_t n) {
  if (pass == 0) {
    if (pcfd_ws_supported_fwd(a, a_ps, lda, w, ldw, b, b_ps, ldb, cj, rows, k, n)) return PCFD_ENGINE_TCGEN05;
    return pcfd_thin_fwd_kind(tin, cj, rows, k, n) >= 0 ? PCFD_ENGINE_THIN : PCFD_ENGINE_FFMA;
  }
  if (pass == 1) {   // a = gzout, b = zin (gzin has the layout of zin)
    if (pcfd_thin_dx_kind(tin, has_gescale ? reinterpret_cast<const float*>(1) : nullptr, cj, rows, k, n) >= 0) return PCFD_ENGINE_THIN;
    return pcfd_ws_supported_dx(a, a_ps, lda, w, ldw, b, b_ps, ldb, b, b_ps, ldb, cj, rows, k, n) ? PCFD_ENGINE_TCGEN05
                                                                                                : PCFD_ENGINE_FFMA;
  }
  if (pass == 2) {   // a = gzout, b = zin
    if (pcfd_small_rows_supported_dw(tin, cj, rows, k, n)) return PCFD_ENGINE_THIN;
    return pcfd_ws_supported_dw(a, a_ps, lda, b, b_ps, ldb, cj, rows, k, n) ? PCFD_ENGINE_TCGEN05 : PCFD_ENGINE_FFMA;
  }
  return -1;
}
