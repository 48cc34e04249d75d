// Exact farthest point sampling for large point sets (n >= ~2k per geometry): bucketed FPS.
//
// torch_cluster.fps (reference models/modules.py:320) is m sequential steps "update every running min-distance with the
// distance to the newest sample, take the arg-max".  Almost all of those updates are no-ops: once a few samples exist, a
// new sample only lowers the min-distance of points NEAR it.  Points are therefore sorted along a Morton curve and cut
// into buckets of 32 consecutive points (a warp's worth) and super-buckets of 32 buckets, each with its bounding box and
// its current maximum (min-distance, index).  For a new sample c, a (super-)bucket whose box lies at least as far from
// c as its current maximum min-distance cannot change -- skipped without touching its points; the others (a handful
// after the first few samples) are updated by one warp each, and the arg-max is a reduction over the per-bucket maxima.
//
// Bit-exactness against the plain algorithm (oracle/pyg_restate.fps; the kernels in points.cu): a processed point is
// updated with the SAME arithmetic (separately rounded fp32 subtract / multiply / add in coordinate order); the box
// distance is computed with the same sequence of rounded operations on |c - box| gaps, and every one of those
// operations is monotone, so  box_distance <= distance(point, c)  holds in fp32 for every point in the box, and
// "box_distance >= bucket maximum" implies fminf(d_p, distance) == d_p for all of them: skipping changes nothing.
// Ties resolve to the lowest ORIGINAL index (the sort permutes storage, not identity).
//
//   fb_bbox_kernel   per geometry: bounding box                       (1 launch)
//   fb_keys_kernel   per point: (geometry << 32 | Morton code), index (1 launch)
//   cub::DeviceRadixSort::SortPairs over all geometries at once       (library sort of the keys; temp storage from the caller)
//   fb_fps_kernel    one CTA (8 warps) per geometry: the sampling loop, one block barrier per sample
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace pcfd {

#ifndef PCFD_FB_WARPS
#define PCFD_FB_WARPS 16
#endif
constexpr int FB_WARPS = PCFD_FB_WARPS;
constexpr int FB_THREADS = FB_WARPS * 32;

template <int DIMS>
__device__ __forceinline__ float fb_sqdist(const float4& p, const float (&c)[3]) {
  const float d0 = __fsub_rn(p.x, c[0]);
  float acc = __fmul_rn(d0, d0);
  const float d1 = __fsub_rn(p.y, c[1]);
  acc = __fadd_rn(acc, __fmul_rn(d1, d1));
  if (DIMS == 3) {
    const float d2 = __fsub_rn(p.z, c[2]);
    acc = __fadd_rn(acc, __fmul_rn(d2, d2));
  }
  return acc;
}

// lower bound (in fp32, see the header) of fb_sqdist(p, c) over all p inside the box [lo, hi]
template <int DIMS>
__device__ __forceinline__ float fb_box_bound(const float* lo, const float* hi, const float (&c)[3]) {
  float acc = 0.f;
#pragma unroll
  for (int d = 0; d < DIMS; ++d) {
    float gap = 0.f;
    if (c[d] < lo[d]) gap = __fsub_rn(lo[d], c[d]);
    else if (c[d] > hi[d]) gap = __fsub_rn(c[d], hi[d]);
    const float sq = __fmul_rn(gap, gap);
    acc = d == 0 ? sq : __fadd_rn(acc, sq);
  }
  return acc;
}

__global__ void __launch_bounds__(256) fb_bbox_kernel(const float* __restrict__ pos, int n, int dims, float* __restrict__ bbox) {
  __shared__ float slo[8][3], shi[8][3];
  const int g = blockIdx.x;
  const float* gp = pos + (size_t)g * n * dims;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int p = threadIdx.x; p < n; p += 256)
    for (int d = 0; d < dims; ++d) { const float v = gp[(size_t)p * dims + d]; lo[d] = fminf(lo[d], v); hi[d] = fmaxf(hi[d], v); }
  for (int d = 0; d < 3; ++d)
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
  if ((threadIdx.x & 31) == 0)
    for (int d = 0; d < 3; ++d) { slo[threadIdx.x >> 5][d] = lo[d]; shi[threadIdx.x >> 5][d] = hi[d]; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float l = slo[0][threadIdx.x], h = shi[0][threadIdx.x];
    for (int w = 1; w < 8; ++w) { l = fminf(l, slo[w][threadIdx.x]); h = fmaxf(h, shi[w][threadIdx.x]); }
    bbox[g * 6 + threadIdx.x] = l;
    bbox[g * 6 + 3 + threadIdx.x] = h;
  }
}

__device__ __forceinline__ uint32_t fb_spread3(uint32_t x) {   // 10 bits -> every third bit
  x &= 0x3ffu;
  x = (x | (x << 16)) & 0x030000ffu;
  x = (x | (x << 8)) & 0x0300f00fu;
  x = (x | (x << 4)) & 0x030c30c3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}
__device__ __forceinline__ uint32_t fb_spread2(uint32_t x) {   // 15 bits -> every second bit
  x &= 0x7fffu;
  x = (x | (x << 8)) & 0x00ff00ffu;
  x = (x | (x << 4)) & 0x0f0f0f0fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}

__global__ void __launch_bounds__(256) fb_keys_kernel(const float* __restrict__ pos, int64_t total, int n, int dims,
                                                      const float* __restrict__ bbox, uint64_t* __restrict__ keys,
                                                      uint32_t* __restrict__ vals) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int g = (int)(i / n);
  const int p = (int)(i - (int64_t)g * n);
  const float* b = bbox + g * 6;
  uint32_t q[3] = {0u, 0u, 0u};
  const float scale = dims == 3 ? 1023.0f : 32767.0f;
  for (int d = 0; d < dims; ++d) {
    const float ext = b[3 + d] - b[d];
    float u = ext > 0.f ? (pos[i * dims + d] - b[d]) / ext : 0.f;
    u = fminf(fmaxf(u, 0.f), 1.f);
    q[d] = (uint32_t)(u * scale);
  }
  const uint32_t code = dims == 3 ? (fb_spread3(q[0]) | (fb_spread3(q[1]) << 1) | (fb_spread3(q[2]) << 2))
                                  : (fb_spread2(q[0]) | (fb_spread2(q[1]) << 1));
  keys[i] = ((uint64_t)(uint32_t)g << 32) | code;
  vals[i] = (uint32_t)p;
}

// Bucket b is owned by (warp = b % FB_WARPS, lane = (b / FB_WARPS) % 32, slot j = b / (32 * FB_WARPS)): neighbours along
// the Morton curve -- the buckets a new sample touches -- land on different warps.  The owner lane keeps the bucket's box
// and its current maximum (distance bits, original index) in REGISTERS; shared memory holds the points and the
// original -> sorted position table (global memory / L2 when a geometry does not fit) and one 64-bit arg-max cell per
// sample that the warps combine with atomicMax (key = distance bits << 32 | ~index: largest distance, lowest index).
template <int DIMS, bool SMEM, int J>
__global__ void __launch_bounds__(FB_THREADS) fb_fps_kernel(const float* __restrict__ pos, int n, int m,
                                                            const uint32_t* __restrict__ order, float4* __restrict__ gpts,
                                                            unsigned* __restrict__ goidx, unsigned* __restrict__ ginv,
                                                            int64_t* __restrict__ idx_out) {
  extern __shared__ __align__(16) unsigned char fb_smem[];
  __shared__ unsigned long long cell[3];
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbk = (n + 31) >> 5, npad = nbk << 5;
  float4* pts;
  unsigned *oidx, *inv;
  if (SMEM) {
    pts = reinterpret_cast<float4*>(fb_smem);
    oidx = reinterpret_cast<unsigned*>(pts + npad);
    inv = oidx + npad;
  } else {
    pts = gpts + (size_t)g * npad;
    oidx = goidx + (size_t)g * npad;
    inv = ginv + (size_t)g * npad;
  }
  // ---- set-up: points in Morton order, distances +inf
  const float* gp = pos + (size_t)g * n * DIMS;
  const uint32_t* ord = order + (size_t)g * n;
  for (int p = tid; p < npad; p += FB_THREADS) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned oi = 0xffffffffu;
    if (p < n) {
      oi = ord[p];
      v.x = gp[(size_t)oi * DIMS];
      v.y = gp[(size_t)oi * DIMS + 1];
      if (DIMS == 3) v.z = gp[(size_t)oi * DIMS + 2];
      v.w = INFINITY;
      inv[oi] = (unsigned)p;
    }
    pts[p] = v;
    oidx[p] = oi;
  }
  if (tid < 3) cell[tid] = 0ULL;
  __syncthreads();
  // ---- boxes and maxima of this lane's buckets
  float blo[J][DIMS], bhi[J][DIMS];
  unsigned kd[J], ki[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    kd[j] = 0u; ki[j] = 0xffffffffu;
#pragma unroll
    for (int d = 0; d < DIMS; ++d) { blo[j][d] = INFINITY; bhi[j][d] = -INFINITY; }
    for (int l = 0; l < 32; ++l) {
      const int b = warp + FB_WARPS * (l + 32 * j);          // uniform over the warp
      if (b >= nbk) break;
      const int p = (b << 5) + lane;
      const float4 v = pts[p];
      const unsigned oi = oidx[p];
      const bool real = oi != 0xffffffffu;
      float lo[3] = {real ? v.x : INFINITY, real ? v.y : INFINITY, real ? v.z : INFINITY};
      float hi[3] = {real ? v.x : -INFINITY, real ? v.y : -INFINITY, real ? v.z : -INFINITY};
#pragma unroll
      for (int d = 0; d < DIMS; ++d)
        for (int o = 16; o > 0; o >>= 1) {
          lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
          hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
      const unsigned wd = __reduce_max_sync(0xffffffffu, __float_as_uint(v.w));
      const unsigned wi = __reduce_min_sync(0xffffffffu, __float_as_uint(v.w) == wd ? oi : 0xffffffffu);
      if (lane == l) {
#pragma unroll
        for (int d = 0; d < DIMS; ++d) { blo[j][d] = lo[d]; bhi[j][d] = hi[d]; }
        kd[j] = wd; ki[j] = wi;
      }
    }
  }

  int curp = (int)inv[0];
  if (tid == 0) idx_out[(size_t)g * m] = (int64_t)g * n;
  for (int s = 1; s < m; ++s) {
    const float4 cv = pts[curp];
    const float c[3] = {cv.x, cv.y, cv.z};
    if (tid == 0) cell[(s + 1) % 3] = 0ULL;       // last read two samples ago
#pragma unroll
    for (int j = 0; j < J; ++j) {
      // a bucket whose box is at least as far from c as its largest min-distance cannot change (empty slots: kd = 0)
      const bool act = !(fb_box_bound<DIMS>(blo[j], bhi[j], c) >= __uint_as_float(kd[j]));
      unsigned mb = __ballot_sync(0xffffffffu, act);
      while (mb) {
        const int l = __ffs(mb) - 1;
        mb &= mb - 1;
        const int p = ((warp + FB_WARPS * (l + 32 * j)) << 5) + lane;
        const float4 v = pts[p];
        const unsigned oi = oidx[p];
        const float d = fminf(v.w, fb_sqdist<DIMS>(v, c));
        if (d != v.w) reinterpret_cast<float*>(pts + p)[3] = d;
        const unsigned wd = __reduce_max_sync(0xffffffffu, __float_as_uint(d));
        const unsigned wi = __reduce_min_sync(0xffffffffu, __float_as_uint(d) == wd ? oi : 0xffffffffu);
        if (lane == l) { kd[j] = wd; ki[j] = wi; }
      }
    }
    // ---- arg-max: this lane's buckets, the warp, the block (largest distance, then lowest original index)
    unsigned bd = kd[0], bi = ki[0];
#pragma unroll
    for (int j = 1; j < J; ++j)
      if (kd[j] > bd || (kd[j] == bd && ki[j] < bi)) { bd = kd[j]; bi = ki[j]; }
    const unsigned wd = __reduce_max_sync(0xffffffffu, bd);
    const unsigned wi = __reduce_min_sync(0xffffffffu, bd == wd ? bi : 0xffffffffu);
    if (lane == 0) atomicMax(&cell[s % 3], ((unsigned long long)wd << 32) | (unsigned long long)(0xffffffffu - wi));
    __syncthreads();
    const unsigned win = 0xffffffffu - (unsigned)(cell[s % 3] & 0xffffffffULL);
    curp = (int)inv[win];
    if (tid == 0) idx_out[(size_t)g * m + s] = (int64_t)g * n + (int64_t)win;
  }
}

struct FbLayout { size_t keys_a, keys_b, vals_a, vals_b, bbox, temp, pts, oidx, inv, total; size_t temp_bytes, smem; bool in_smem; int end_bit, j; };

static inline size_t fb_align(size_t x) { return (x + 255) & ~(size_t)255; }

static FbLayout fb_layout(int n_geom, int n) {
  FbLayout L{};
  const size_t N = (size_t)n_geom * n;
  const int nbk = (n + 31) / 32, npad = nbk * 32;
  int j = 1;
  while (j * 32 * FB_WARPS < nbk) j *= 2;
  L.j = j;
  int gbits = 0;
  while ((1 << gbits) < n_geom) ++gbits;
  L.end_bit = 32 + gbits;
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)N, 0, L.end_bit, (cudaStream_t)0);
  L.temp_bytes = temp;
  const size_t with_pts = (size_t)npad * 24;
  L.in_smem = with_pts <= 200 * 1024;
  L.smem = L.in_smem ? with_pts : 0;
  size_t o = 0;
  L.keys_a = o; o += fb_align(N * 8);
  L.keys_b = o; o += fb_align(N * 8);
  L.vals_a = o; o += fb_align(N * 4);
  L.vals_b = o; o += fb_align(N * 4);
  L.bbox = o; o += fb_align((size_t)n_geom * 6 * 4);
  L.temp = o; o += fb_align(temp);
  L.pts = o; o += L.in_smem ? 0 : fb_align((size_t)n_geom * npad * 16);
  L.oidx = o; o += L.in_smem ? 0 : fb_align((size_t)n_geom * npad * 4);
  L.inv = o; o += L.in_smem ? 0 : fb_align((size_t)n_geom * npad * 4);
  L.total = o;
  return L;
}

}  // namespace pcfd

using namespace pcfd;

extern "C" int pcfd_fps_bucket_supported(int32_t n_geom, int32_t n, int32_t dims) {
  if (n_geom <= 0 || n < 64 || (dims != 2 && dims != 3)) return 0;
  if ((int64_t)n_geom * n >= ((int64_t)1 << 31)) return 0;
  return fb_layout(n_geom, n).j <= 16;      // up to 131072 points per geometry
}

extern "C" size_t pcfd_fps_bucket_workspace_bytes(int32_t n_geom, int32_t n, int32_t dims) {
  if (!pcfd_fps_bucket_supported(n_geom, n, dims)) return 0;
  return fb_layout(n_geom, n).total;
}

extern "C" int pcfd_fps_bucket(const float* pos, int32_t n_geom, int32_t n, int32_t dims, int32_t m, int64_t* idx_out,
                               void* workspace, size_t workspace_bytes, void* stream) {
  if (!pos || !idx_out || m <= 0 || m > n || !pcfd_fps_bucket_supported(n_geom, n, dims)) return PCFD_ERR_ARG;
  const FbLayout L = fb_layout(n_geom, n);
  if (!workspace || workspace_bytes < L.total) return PCFD_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return PCFD_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
  uint64_t* keys_a = reinterpret_cast<uint64_t*>(w + L.keys_a);
  uint64_t* keys_b = reinterpret_cast<uint64_t*>(w + L.keys_b);
  uint32_t* vals_a = reinterpret_cast<uint32_t*>(w + L.vals_a);
  uint32_t* vals_b = reinterpret_cast<uint32_t*>(w + L.vals_b);
  float* bbox = reinterpret_cast<float*>(w + L.bbox);
  const int64_t N = (int64_t)n_geom * n;
  fb_bbox_kernel<<<n_geom, 256, 0, st>>>(pos, n, dims, bbox);
  PCFD_CHECK_LAUNCH();
  fb_keys_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(pos, N, n, dims, bbox, keys_a, vals_a);
  PCFD_CHECK_LAUNCH();
  size_t temp = L.temp_bytes;
  const cudaError_t se = cub::DeviceRadixSort::SortPairs(w + L.temp, temp, (const uint64_t*)keys_a, keys_b, (const uint32_t*)vals_a,
                                                         vals_b, (int)N, 0, L.end_bit, st);
  if (se != cudaSuccess) return PCFD_ERR_CUDA + (int)se;
  float4* gpts = L.in_smem ? nullptr : reinterpret_cast<float4*>(w + L.pts);
  unsigned* goidx = L.in_smem ? nullptr : reinterpret_cast<unsigned*>(w + L.oidx);
  unsigned* ginv = L.in_smem ? nullptr : reinterpret_cast<unsigned*>(w + L.inv);
  cudaError_t e;
#define PCFD_FB(D_, S_, J_)                                                                                          \
  {                                                                                                                  \
    e = cudaFuncSetAttribute(fb_fps_kernel<D_, S_, J_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);   \
    if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;                                                             \
    fb_fps_kernel<D_, S_, J_><<<n_geom, FB_THREADS, L.smem, st>>>(pos, n, m, vals_b, gpts, goidx, ginv, idx_out);         \
  }
#define PCFD_FB_J(D_, S_)                                                                                            \
  switch (L.j) {                                                                                                     \
    case 1: PCFD_FB(D_, S_, 1) break;                                                                                \
    case 2: PCFD_FB(D_, S_, 2) break;                                                                                \
    case 4: PCFD_FB(D_, S_, 4) break;                                                                                \
    case 8: PCFD_FB(D_, S_, 8) break;                                                                                \
    default: PCFD_FB(D_, S_, 16) break;                                                                              \
  }
  if (dims == 2) { if (L.in_smem) PCFD_FB_J(2, true) else PCFD_FB_J(2, false) }
  else { if (L.in_smem) PCFD_FB_J(3, true) else PCFD_FB_J(3, false) }
#undef PCFD_FB_J
#undef PCFD_FB
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}
