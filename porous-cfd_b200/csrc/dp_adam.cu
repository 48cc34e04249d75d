// Data-parallel optimizer tail in ONE kernel over NVLink / NVSwitch peer memory: gradient reduce-scatter, Adam on this
// rank's slice, all-gather of the updated parameters -- replacing  ncclAllReduce(flat_grad) + adam_kernel  (reference:
// Lightning DDP's bucketed all-reduce + torch.optim.Adam, common/training.py:66-71,83, models/pipn/pipn_foam.py:102-105).
//
// Every rank's flat gradient, flat parameters and a small flag array live in symmetric memory (same size on every rank,
// mapped into every peer; the host side exchanges the handles, common/training.py).  With W ranks and n parameters:
//
//   barrier A   CTA b of every rank tells CTA b of every peer "my gradients are complete" and waits for theirs
//   slice       rank r owns float4 units [r*per, (r+1)*per): it reads the SUM of all ranks' gradients there --
//               one multimem.ld_reduce.add.v4.f32 on the multicast address (the NVSwitch adds the W copies in flight),
//               or W peer loads added in rank order where multicast is not available -- applies Adam with its local
//               moments, and writes the new parameters to ALL ranks (multimem.st on the multicast address / W peer stores)
//   barrier B   "my stores have landed" to every peer; a rank's kernel ends when all its CTAs have heard from all peers
//
// Traffic per rank: n/W * 4 B * (W loads + W stores) instead of the 2 * (W-1)/W * n * 4 B of a ring plus a second pass over
// n for Adam; one launch, no NCCL protocol latency (measured: RING_LL, 39 us at W = 2 for 3.4 MB).  The moments are sharded
// for free (a rank only ever touches its slice).  Flags carry a launch counter (monotonic, never reset), so the kernel is
// re-entrant across CUDA-graph replays; a spin that exceeds ~2 s sets epoch[2] and gives up instead of hanging the GPU.
#include "common.cuh"

namespace pcfd {

constexpr int DP_MAX_WORLD = 16;
constexpr int DP_CTAS = 64;
constexpr int DP_THREADS = 256;

struct DpArgs {
  float* grad[DP_MAX_WORLD];
  float* param[DP_MAX_WORLD];
  uint32_t* flags[DP_MAX_WORLD];
  float* grad_mc; float* param_mc;
  float* m; float* v;
  const int64_t* step; const float* lr;
  int32_t* epoch;            // [0] launch counter, [1] ticket, [2] error
  float b1, b2, eps, grad_scale;
  int64_t n;
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys_v4(const float* p) {      // peer memory: not through this SM's L1
  float4 r;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// CTA b of this rank <-> CTA b of every peer.  flags layout on every rank: [phase][cta][source rank]
__device__ __forceinline__ void cross_rank_barrier(const DpArgs& a, int phase, uint32_t val) {
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < a.world) {
    const size_t slot = ((size_t)phase * gridDim.x + blockIdx.x) * DP_MAX_WORLD;
    st_release_sys(a.flags[threadIdx.x] + slot + a.rank, val);
    const uint32_t* mine = a.flags[a.rank] + slot + threadIdx.x;
    long long spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - val) < 0) {
      if (++spins > (1LL << 27)) { a.epoch[2] = 1 + phase; break; }      // ~2 s: report, do not hang
      __nanosleep(20);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(DP_THREADS) dp_adam_kernel(DpArgs a) {
  __shared__ int is_last;
  const uint32_t ep = (uint32_t) * reinterpret_cast<volatile int32_t*>(a.epoch);
  cross_rank_barrier(a, 0, ep + 1);

  const int64_t n4 = (a.n + 3) / 4;
  const int64_t per = (n4 + a.world - 1) / a.world;
  const int64_t lo = (int64_t)a.rank * per;
  const int64_t hi = lo + per < n4 ? lo + per : n4;
  const float t = (float)(*a.step);
  const float bc1 = 1.0f - powf(a.b1, t), bc2 = 1.0f - powf(a.b2, t);
  const float step_size = *a.lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    float4 g;
    if (a.grad_mc != nullptr) {
      g = multimem_ld_reduce_add(a.grad_mc + 4 * i);
    } else {
      g = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < a.world; ++r) {
        const float4 q = ld_sys_v4(a.grad[r] + 4 * i);
        g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
      }
    }
    float gr[4] = {g.x * a.grad_scale, g.y * a.grad_scale, g.z * a.grad_scale, g.w * a.grad_scale};
    float4 m4 = *reinterpret_cast<float4*>(a.m + 4 * i), v4 = *reinterpret_cast<float4*>(a.v + 4 * i);
    float4 p4 = *reinterpret_cast<float4*>(a.param[a.rank] + 4 * i);
    float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      mm[e] = a.b1 * mm[e] + (1.0f - a.b1) * gr[e];
      vv[e] = a.b2 * vv[e] + (1.0f - a.b2) * gr[e] * gr[e];
      pp[e] -= step_size * mm[e] / (sqrtf(vv[e]) * inv_sqrt_bc2 + a.eps);
    }
    *reinterpret_cast<float4*>(a.m + 4 * i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(a.v + 4 * i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    const float4 pn = make_float4(pp[0], pp[1], pp[2], pp[3]);
    if (a.param_mc != nullptr) {
      multimem_st(a.param_mc + 4 * i, pn);
    } else {
      for (int r = 0; r < a.world; ++r) *reinterpret_cast<float4*>(a.param[r] + 4 * i) = pn;
    }
  }

  cross_rank_barrier(a, 1, ep + 1);
  // the CTA that finishes last advances the launch counter (every CTA has read it by then)
  if (threadIdx.x == 0) is_last = atomicAdd(a.epoch + 1, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    a.epoch[1] = 0;
    __threadfence();
    a.epoch[0] = (int32_t)(ep + 1);
  }
}

__global__ void dp_adam_advance_kernel(int64_t* step) { *step += 1; }

}  // namespace pcfd

using namespace pcfd;

// number of uint32 flags every rank's symmetric flag buffer must hold (zero before the first call)
extern "C" int32_t pcfd_dp_flags_len() { return 2 * DP_CTAS * DP_MAX_WORLD; }

extern "C" int pcfd_dp_adam_step(const pcfd_dp_peers_t* peers, float* exp_avg, float* exp_avg_sq, int64_t* step,
                                 const float* lr, float beta1, float beta2, float eps, float grad_scale, int64_t n,
                                 int32_t* epoch, void* stream) {
  if (!peers || !exp_avg || !exp_avg_sq || !step || !lr || !epoch || n <= 0) return PCFD_ERR_ARG;
  if (peers->world < 1 || peers->world > DP_MAX_WORLD || peers->rank < 0 || peers->rank >= peers->world) return PCFD_ERR_ARG;
  DpArgs a;
  for (int r = 0; r < DP_MAX_WORLD; ++r) {
    a.grad[r] = r < peers->world ? reinterpret_cast<float*>(peers->grad[r]) : nullptr;
    a.param[r] = r < peers->world ? reinterpret_cast<float*>(peers->param[r]) : nullptr;
    a.flags[r] = r < peers->world ? reinterpret_cast<uint32_t*>(peers->flags[r]) : nullptr;
    if (r < peers->world && (!a.grad[r] || !a.param[r] || !a.flags[r])) return PCFD_ERR_ARG;
    if (r < peers->world && ((reinterpret_cast<uintptr_t>(a.grad[r]) | reinterpret_cast<uintptr_t>(a.param[r])) & 15)) return PCFD_ERR_ARG;
  }
  a.grad_mc = reinterpret_cast<float*>(peers->grad_mc);
  a.param_mc = reinterpret_cast<float*>(peers->param_mc);
  if ((a.grad_mc == nullptr) != (a.param_mc == nullptr)) return PCFD_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) return PCFD_ERR_ARG;
  a.m = exp_avg; a.v = exp_avg_sq; a.step = step; a.lr = lr; a.epoch = epoch;
  a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.grad_scale = grad_scale; a.n = n;
  a.rank = peers->rank; a.world = peers->world;
  cudaStream_t st = (cudaStream_t)stream;
  dp_adam_advance_kernel<<<1, 1, 0, st>>>(step);
  PCFD_CHECK_LAUNCH();
  dp_adam_kernel<<<DP_CTAS, DP_THREADS, 0, st>>>(a);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}
