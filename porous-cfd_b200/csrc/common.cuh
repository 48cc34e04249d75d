// Shared device helpers for libpcfd_sm100: jet activation algebra, dropout hash, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdlib>
#include <utility>

#include "../../include/pcfd.h"

#define PCFD_CHECK_LAUNCH()                                   \
  do {                                                        \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) return PCFD_ERR_CUDA + (int)e__;  \
  } while (0)

namespace pcfd {

// PCFD_PDL=0 launches everything fully serialised
inline int pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("PCFD_PDL"); on = e ? atoi(e) : 1; }
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-(kernel, device) attribute: set it the first time a device sees the
// kernel (or needs more than before).  The record is an idempotent cache, one atomic per device: safe from several host
// threads and with several devices in one process (a plain `static bool configured` is neither).
template <auto Kernel>
inline cudaError_t ensure_dyn_smem(int bytes) {
  static std::atomic<int> have[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cudaErrorInvalidDevice;
  std::atomic<int>& h = have[dev & 63];
  if (h.load(std::memory_order_acquire) >= bytes) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) {
    int cur = h.load(std::memory_order_relaxed);
    while (cur < bytes && !h.compare_exchange_weak(cur, bytes, std::memory_order_release)) {}
  }
  return e;
}

// ---- programmatic dependent launch -----------------------------------------------------------------
// A kernel launched through launch_pdl may become resident while its predecessor in the stream is still running: its
// prologue (barrier init, TMEM allocation, tensor-map prefetch) overlaps the predecessor's tail, and griddep_wait()
// holds it until the predecessor has completed and its writes are visible.  Nothing before griddep_wait() may touch
// global memory.  griddep_launch_dependents() lets the NEXT kernel of the stream be scheduled as SMs free up.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }


// ---- jet layout: cj -> (spatial dims D, derivative order) ---------------------------------
template <int CJ> struct JetShape;
template <> struct JetShape<1> { static constexpr int D = 0, ORDER = 0; };
template <> struct JetShape<3> { static constexpr int D = 2, ORDER = 1; };
template <> struct JetShape<4> { static constexpr int D = 3, ORDER = 1; };
template <> struct JetShape<5> { static constexpr int D = 2, ORDER = 2; };
template <> struct JetShape<7> { static constexpr int D = 3, ORDER = 2; };

__host__ __device__ inline bool valid_cj(int cj) { return cj == 1 || cj == 3 || cj == 4 || cj == 5 || cj == 7; }

// ---- activation and its first three derivatives -------------------------------------------
struct ActD { float f0, f1, f2, f3; };

// sigmoid / tanh on the SFU (ex2.approx + rcp): ~1e-6 relative error for |z| <= 10, two orders of
// magnitude inside the 1e-4 parity budget, and ~4x fewer instructions than expf + IEEE division --
// the input transform is on the critical path of the tensor-core kernels' operand staging.
// Raw MUFU forms: ex2 saturates to +inf / 0 and rcp(+inf) = 0, so no range fix-up is needed (the __expf / __fdividef
// wrappers add 3-4 instructions of it per call, which matters: the transform warps are issue-bound).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_sigmoid(float z) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * z)); }
__device__ __forceinline__ float fast_tanh(float z) {
  const float z2 = z * z;
  if (z2 < 0.01f) return z * (1.0f + z2 * (-0.33333333f + z2 * (0.13333333f - z2 * 0.053968254f)));
  return 1.0f - 2.0f * rcp_approx(1.0f + ex2_approx(2.8853900817779268f * z));
}

template <bool NEED3>
__device__ __forceinline__ ActD act_derivs(int act, float z) {
  ActD r;
  if (act == PCFD_ACT_SILU) {
    // silu(z) = z*s, s = sigmoid(z); t = s(1-s)
    float s = fast_sigmoid(z);
    float t = s * (1.0f - s);
    float q = 1.0f - 2.0f * s;
    r.f0 = z * s;
    r.f1 = s + z * t;
    r.f2 = t * (2.0f + z * q);
    r.f3 = NEED3 ? t * (q * (3.0f + z * q) - 2.0f * z * t) : 0.0f;
  } else if (act == PCFD_ACT_TANH) {
    float t = fast_tanh(z);
    float d = 1.0f - t * t;
    r.f0 = t;
    r.f1 = d;
    r.f2 = -2.0f * t * d;
    r.f3 = NEED3 ? -2.0f * d * (1.0f - 3.0f * t * t) : 0.0f;
  } else {
    r.f0 = z; r.f1 = 1.0f; r.f2 = 0.0f; r.f3 = 0.0f;
  }
  return r;
}

// compile-time activation: the tensor-core engines dispatch once per tile on the (kernel-uniform) activation
template <int ACT, bool NEED2, bool NEED3>
__device__ __forceinline__ ActD act_derivs_t(float z) {
  ActD r;
  if (ACT == PCFD_ACT_SILU) {
    const float s = fast_sigmoid(z);
    const float t = s * (1.0f - s);
    const float q = 1.0f - 2.0f * s;
    r.f0 = z * s;
    r.f1 = s + z * t;
    r.f2 = NEED2 ? t * (2.0f + z * q) : 0.0f;
    r.f3 = NEED3 ? t * (q * (3.0f + z * q) - 2.0f * z * t) : 0.0f;
  } else if (ACT == PCFD_ACT_TANH) {
    const float t = fast_tanh(z);
    const float d = 1.0f - t * t;
    r.f0 = t;
    r.f1 = d;
    r.f2 = NEED2 ? -2.0f * t * d : 0.0f;
    r.f3 = NEED3 ? -2.0f * d * (1.0f - 3.0f * t * t) : 0.0f;
  } else {
    r.f0 = z; r.f1 = 1.0f; r.f2 = 0.0f; r.f3 = 0.0f;
  }
  return r;
}

__device__ __forceinline__ float act_value(int act, float z) {
  if (act == PCFD_ACT_SILU) return z * fast_sigmoid(z);
  if (act == PCFD_ACT_TANH) return fast_tanh(z);
  return z;
}
__device__ __forceinline__ float act_d1(int act, float z) {
  if (act == PCFD_ACT_SILU) { float s = fast_sigmoid(z); return s + z * s * (1.0f - s); }
  if (act == PCFD_ACT_TANH) { float t = fast_tanh(z); return 1.0f - t * t; }
  return 1.0f;
}

// ---- dropout: counter-based hash of (seed, salt, row, col) --------------------------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
// multiplier applied to element (row, col): 0 (dropped) or 1/(1-p).  The hash is staged (seed+salt, then
// row, then column) so that callers can hoist the first two levels out of their column loops.
__host__ __device__ __forceinline__ uint32_t dropout_seed_hash(uint64_t seed, uint32_t salt) {
  return mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + salt));
}
__host__ __device__ __forceinline__ uint32_t dropout_row_hash(uint32_t hseed, int64_t row) {
  return mix32(hseed ^ (uint32_t)row);
}
__device__ __forceinline__ float dropout_from_row(uint32_t hrow, int64_t row, int col, float p, float inv_keep) {
  const uint32_t h = mix32(hrow + 0x9e3779b9U * (uint32_t)col + (uint32_t)((uint64_t)row >> 32));
  // u = (h >> 8) / 2^24 is uniform in [0,1): keep when u >= p.  Both sides scaled by 2^24 are exact, so the test is the
  // integer comparison (h >> 8) >= ceil(p * 2^24); the threshold is loop-invariant and the int->float conversion (an
  // XU-pipe instruction, like the activation's MUFU ops) goes away
  const uint32_t thr = (uint32_t)ceilf(p * 16777216.0f);
  return (h >> 8) >= thr ? inv_keep : 0.0f;
}
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint32_t salt, int64_t row, int col, float p, float inv_keep) {
  return dropout_from_row(dropout_row_hash(dropout_seed_hash(seed, salt), row), row, col, p, inv_keep);
}

// Device-side view of pcfd_intrans_t (copied by value into kernel parameters).
struct InTrans {
  int act;
  int act_cols;        // resolved: number of leading columns the transform applies to
  const float* escale;
  int ldescale;
  float drop_p;
  float inv_keep;
  const uint64_t* seed_dev;
  uint32_t salt;
};

inline InTrans make_intrans(const pcfd_intrans_t* t, int k) {
  InTrans r;
  if (t == nullptr) {
    r.act = PCFD_ACT_NONE; r.act_cols = k; r.escale = nullptr; r.ldescale = 0; r.drop_p = 0.f; r.inv_keep = 1.f;
    r.seed_dev = nullptr; r.salt = 0;
    return r;
  }
  r.act = t->act;
  r.act_cols = (t->act_cols <= 0 || t->act_cols > k) ? k : t->act_cols;
  r.escale = t->escale; r.ldescale = t->ldescale;
  r.drop_p = t->drop_p;
  r.inv_keep = t->drop_p > 0.f ? 1.0f / (1.0f - t->drop_p) : 1.0f;
  r.seed_dev = t->seed_dev; r.salt = t->salt;
  return r;
}

// scale s = dropout mask * escale for input element (row, col); `m` receives the mask part alone
// geometry index of a row (32-bit division: 64-bit integer division costs > 100 instructions)
__device__ __forceinline__ int64_t geom_of(int64_t row, int64_t rows_per_geom) {
  return rows_per_geom > 0 ? (int64_t)((uint32_t)row / (uint32_t)rows_per_geom) : 0;
}

__device__ __forceinline__ float in_scale(const InTrans& t, uint64_t seed, int64_t row, int64_t geom, int col, float& m) {
  m = 1.0f;
  if (t.drop_p > 0.0f) m = dropout_scale(seed, t.salt, row, col, t.drop_p, t.inv_keep);
  float s = m;
  if (t.escale != nullptr) s *= __ldg(t.escale + geom * t.ldescale + col);
  return s;
}

// Forward input transform of one element, all channels.  z[] holds the CJ pre-activation
// channels on entry and the transformed channels on exit.
template <int CJ>
__device__ __forceinline__ void jet_act_fwd(int act, float s, float (&z)[CJ]) {
  constexpr int D = JetShape<CJ>::D;
  constexpr int ORDER = JetShape<CJ>::ORDER;
  if (act == PCFD_ACT_NONE) {
#pragma unroll
    for (int c = 0; c < CJ; ++c) z[c] *= s;
    return;
  }
  ActD a = act_derivs<false>(act, z[0]);
  z[0] = a.f0 * s;
  if (ORDER >= 1) {
    float f1s = a.f1 * s, f2s = a.f2 * s;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      float zk = z[1 + k];
      if (ORDER >= 2) z[1 + D + k] = f2s * zk * zk + f1s * z[1 + D + k];
      z[1 + k] = f1s * zk;
    }
  }
}

// jet_act_fwd with the activation fixed at compile time
template <int CJ, int ACT>
__device__ __forceinline__ void jet_act_fwd_t(float s, float (&z)[CJ]) {
  constexpr int D = JetShape<CJ>::D;
  constexpr int ORDER = JetShape<CJ>::ORDER;
  if (ACT == PCFD_ACT_NONE) {
#pragma unroll
    for (int c = 0; c < CJ; ++c) z[c] *= s;
    return;
  }
  const ActD a = act_derivs_t<ACT, (ORDER >= 2), false>(z[0]);
  z[0] = a.f0 * s;
  if (ORDER >= 1) {
    const float f1s = a.f1 * s, f2s = a.f2 * s;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const float zk = z[1 + k];
      if (ORDER >= 2) z[1 + D + k] = f2s * zk * zk + f1s * z[1 + D + k];
      z[1 + k] = f1s * zk;
    }
  }
}

// Reverse of jet_act_fwd.  On entry g[] = gradient wrt the transformed channels, z[] = the
// pre-activation channels.  On exit g[] = gradient wrt z[].  Returns sum_c g_in[c] * (a[c]/escale)
// (the contribution to the escale gradient) where m is the dropout part of s.
template <int CJ>
__device__ __forceinline__ float jet_act_bwd(int act, float s, float m, const float (&z)[CJ], float (&g)[CJ]) {
  constexpr int D = JetShape<CJ>::D;
  constexpr int ORDER = JetShape<CJ>::ORDER;
  float ge = 0.0f;
  if (act == PCFD_ACT_NONE) {
#pragma unroll
    for (int c = 0; c < CJ; ++c) { ge += g[c] * z[c] * m; g[c] *= s; }
    return ge;
  }
  ActD a = act_derivs<(ORDER >= 2)>(act, z[0]);
  ge = g[0] * a.f0;
  float g0 = g[0] * s * a.f1;
  if (ORDER >= 1) {
#pragma unroll
    for (int k = 0; k < D; ++k) {
      float zk = z[1 + k];
      float gk = g[1 + k];
      ge += gk * a.f1 * zk;
      float gks = gk * s;
      g0 += gks * a.f2 * zk;
      float gzk = gks * a.f1;
      if (ORDER >= 2) {
        float zkk = z[1 + D + k];
        float gkk = g[1 + D + k];
        ge += gkk * (a.f2 * zk * zk + a.f1 * zkk);
        float gkks = gkk * s;
        g0 += gkks * (a.f3 * zk * zk + a.f2 * zkk);
        gzk += gkks * 2.0f * a.f2 * zk;
        g[1 + D + k] = gkks * a.f1;
      }
      g[1 + k] = gzk;
    }
  }
  g[0] = g0;
  return ge * m;
}

// jet_act_bwd with the activation fixed at compile time
template <int CJ, int ACT>
__device__ __forceinline__ float jet_act_bwd_t(float s, float m, const float (&z)[CJ], float (&g)[CJ]) {
  constexpr int D = JetShape<CJ>::D;
  constexpr int ORDER = JetShape<CJ>::ORDER;
  float ge = 0.0f;
  if (ACT == PCFD_ACT_NONE) {
#pragma unroll
    for (int c = 0; c < CJ; ++c) { ge += g[c] * z[c] * m; g[c] *= s; }
    return ge;
  }
  const ActD a = act_derivs_t<ACT, (ORDER >= 1), (ORDER >= 2)>(z[0]);
  ge = g[0] * a.f0;
  float g0 = g[0] * s * a.f1;
  if (ORDER >= 1) {
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const float zk = z[1 + k];
      const float gk = g[1 + k];
      ge += gk * a.f1 * zk;
      const float gks = gk * s;
      g0 += gks * a.f2 * zk;
      float gzk = gks * a.f1;
      if (ORDER >= 2) {
        const float zkk = z[1 + D + k];
        const float gkk = g[1 + D + k];
        ge += gkk * (a.f2 * zk * zk + a.f1 * zkk);
        const float gkks = gkk * s;
        g0 += gkks * (a.f3 * zk * zk + a.f2 * zkk);
        gzk += gkks * 2.0f * a.f2 * zk;
        g[1 + D + k] = gkks * a.f1;
      }
      g[1 + k] = gzk;
    }
  }
  g[0] = g0;
  return ge * m;
}

inline int check_sm100() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return PCFD_ERR_ARCH;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return PCFD_ERR_ARCH;
  return major == 10 ? PCFD_OK : PCFD_ERR_ARCH;
}

}  // namespace pcfd
