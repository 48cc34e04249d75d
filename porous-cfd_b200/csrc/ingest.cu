// Batch ingestion on the device: the per-case features the reference's FoamDataset adds on the host
// while loading a case (dataset/foam_dataset.py:360-395) and the collation of resident geometries into a batch
// (dataset/foam_dataset.py:83-90), so that a dataset that fits in HBM never crosses PCIe again after it was loaded.
//
//   sdf_min_kernel      unsigned distance of every point of a geometry to its nearest boundary point (scipy cdist + min);
//                       boundary points staged through shared memory in tiles, one thread per query point; compute-bound
//                       (n * n_boundary distance evaluations per geometry), per-geometry maximum by atomicMax on the bits
//   sdf_finish_kernel   divide by the geometry's largest distance, sign from the region flag of the internal points
//   one_hot_kernel      boundaryId columns: zeros for internal rows, one-hot class of the boundary rows
//   gather_blocks       collate_fn: out[i] = src[ids[i]] for fixed-size blocks (16-byte copies); HBM-bound
#include "common.cuh"

namespace pcfd {

constexpr int SDF_TILE = 1024;

template <int D>
__global__ void __launch_bounds__(256) sdf_min_kernel(const float* __restrict__ data, int ld, int pos_col, int n_points,
                                                      int n_internal, const float* __restrict__ coord_scale,
                                                      float* __restrict__ dist, unsigned* __restrict__ gmax) {
  __shared__ float tgt[D][SDF_TILE];
  __shared__ float wmax[8];
  const int64_t g = blockIdx.y;
  const float* base = data + g * (int64_t)n_points * ld + pos_col;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const bool valid = p < n_points;
  float sc[D], x[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    sc[k] = coord_scale != nullptr ? __ldg(coord_scale + k) : 1.0f;
    x[k] = valid ? __ldg(base + (int64_t)p * ld + k) * sc[k] : 0.0f;
  }
  float best = INFINITY;
  for (int t0 = n_internal; t0 < n_points; t0 += SDF_TILE) {
    const int cnt = min(SDF_TILE, n_points - t0);
    for (int i = threadIdx.x; i < cnt; i += 256) {
#pragma unroll
      for (int k = 0; k < D; ++k) tgt[k][i] = __ldg(base + (int64_t)(t0 + i) * ld + k) * sc[k];
    }
    __syncthreads();
    if (valid) {
#pragma unroll 4
      for (int i = 0; i < cnt; ++i) {
        float d2 = 0.0f;
#pragma unroll
        for (int k = 0; k < D; ++k) { const float d = x[k] - tgt[k][i]; d2 = fmaf(d, d, d2); }
        best = fminf(best, d2);
      }
    }
    __syncthreads();
  }
  const float d = valid ? sqrtf(best) : 0.0f;
  if (valid) dist[g * n_points + p] = d;
  // largest distance of the geometry: warp -> block -> one atomic (distances are >= 0: the bit pattern orders like the value)
  float m = d;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = wmax[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) b = fmaxf(b, wmax[w]);
    atomicMax(gmax + g, __float_as_uint(b));
  }
}

__global__ void __launch_bounds__(256) sdf_finish_kernel(float* __restrict__ data, int ld, int sdf_col, int region_col,
                                                         int n_points, int n_internal, const float* __restrict__ dist,
                                                         const unsigned* __restrict__ gmax) {
  const int64_t g = blockIdx.y;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= n_points) return;
  float* row = data + (g * n_points + p) * (int64_t)ld;
  const float scale = __uint_as_float(__ldg(gmax + g));
  // internal points: (0.5 - cellToRegion) * 2 = +1 in the fluid, -1 in the porous region; boundary points are positive
  const float sign = (p < n_internal && region_col >= 0) ? (0.5f - row[region_col]) * 2.0f : 1.0f;
  row[sdf_col] = dist[g * n_points + p] / scale * sign;
}

__global__ void __launch_bounds__(256) one_hot_kernel(float* __restrict__ data, int ld, int col0, int n_classes,
                                                      int n_points, int n_internal, const int32_t* __restrict__ cls) {
  const int64_t g = blockIdx.y;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= n_points) return;
  float* row = data + (g * n_points + p) * (int64_t)ld + col0;
  const int c = p >= n_internal ? __ldg(cls + g * (n_points - n_internal) + (p - n_internal)) : -1;
  for (int k = 0; k < n_classes; ++k) row[k] = k == c ? 1.0f : 0.0f;
}

template <typename T>
__global__ void __launch_bounds__(256) gather_blocks_kernel(const T* __restrict__ src, int64_t block_elems,
                                                            const int64_t* __restrict__ ids, T* __restrict__ dst) {
  const int64_t b = blockIdx.y;
  const T* s = src + __ldg(ids + b) * block_elems;
  T* d = dst + b * block_elems;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < block_elems; i += (int64_t)gridDim.x * 256) d[i] = __ldg(s + i);
}

// several tensors of the same dataset (data + every sub-domain's ids) in one launch: blockIdx.z = tensor
struct GatherTable {
  const void* src[PCFD_GATHER_MAX_TENSORS];
  void* dst[PCFD_GATHER_MAX_TENSORS];
  int64_t elems[PCFD_GATHER_MAX_TENSORS];      // per block, in units of 16 bytes (vec) or 4 bytes
  int vec[PCFD_GATHER_MAX_TENSORS];
};

__global__ void __launch_bounds__(256) gather_blocks_multi_kernel(const __grid_constant__ GatherTable t,
                                                                  const int64_t* __restrict__ ids) {
  const int k = blockIdx.z;
  const int64_t b = blockIdx.y, id = __ldg(ids + b), n = t.elems[k];
  if (t.vec[k]) {
    const uint4* s = reinterpret_cast<const uint4*>(t.src[k]) + id * n;
    uint4* d = reinterpret_cast<uint4*>(t.dst[k]) + b * n;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) d[i] = __ldg(s + i);
  } else {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(t.src[k]) + id * n;
    uint32_t* d = reinterpret_cast<uint32_t*>(t.dst[k]) + b * n;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) d[i] = __ldg(s + i);
  }
}

}  // namespace pcfd

using namespace pcfd;

extern "C" int pcfd_sdf_feature(float* data, int32_t n_geom, int32_t n_points, int32_t f, int32_t n_internal,
                                int32_t pos_col, int32_t dims, int32_t region_col, int32_t sdf_col,
                                const float* coord_scale, float* scratch, void* stream) {
  if (!data || !scratch || n_geom <= 0 || n_points <= 0 || f <= 0) return PCFD_ERR_ARG;
  if (n_internal < 0 || n_internal >= n_points) return PCFD_ERR_ARG;       // at least one boundary point
  if (dims < 2 || dims > 3 || pos_col < 0 || pos_col + dims > f || sdf_col < 0 || sdf_col >= f || region_col >= f)
    return PCFD_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  float* dist = scratch;
  unsigned* gmax = reinterpret_cast<unsigned*>(scratch + (int64_t)n_geom * n_points);
  if (cudaMemsetAsync(gmax, 0, sizeof(unsigned) * n_geom, st) != cudaSuccess) return PCFD_ERR_CUDA;
  for (int32_t g0 = 0; g0 < n_geom; g0 += 65535) {             // grid.y limit
    const int32_t ng = n_geom - g0 < 65535 ? n_geom - g0 : 65535;
    float* d0 = data + (int64_t)g0 * n_points * f;
    dim3 grid((unsigned)((n_points + 255) / 256), (unsigned)ng);
    if (dims == 2) sdf_min_kernel<2><<<grid, 256, 0, st>>>(d0, f, pos_col, n_points, n_internal, coord_scale, dist + (int64_t)g0 * n_points, gmax + g0);
    else sdf_min_kernel<3><<<grid, 256, 0, st>>>(d0, f, pos_col, n_points, n_internal, coord_scale, dist + (int64_t)g0 * n_points, gmax + g0);
    PCFD_CHECK_LAUNCH();
    sdf_finish_kernel<<<grid, 256, 0, st>>>(d0, f, sdf_col, region_col, n_points, n_internal, dist + (int64_t)g0 * n_points, gmax + g0);
    PCFD_CHECK_LAUNCH();
  }
  return PCFD_OK;
}

extern "C" size_t pcfd_sdf_scratch_bytes(int32_t n_geom, int32_t n_points) {
  return ((size_t)n_geom * n_points + (size_t)n_geom) * sizeof(float);
}

extern "C" int pcfd_boundary_one_hot(float* data, int32_t n_geom, int32_t n_points, int32_t f, int32_t n_internal,
                                     const int32_t* boundary_class, int32_t n_classes, int32_t col0, void* stream) {
  if (!data || !boundary_class || n_geom <= 0 || n_points <= 0 || n_internal < 0 || n_internal > n_points) return PCFD_ERR_ARG;
  if (n_classes <= 0 || col0 < 0 || col0 + n_classes > f) return PCFD_ERR_ARG;
  for (int32_t g0 = 0; g0 < n_geom; g0 += 65535) {             // grid.y limit
    const int32_t ng = n_geom - g0 < 65535 ? n_geom - g0 : 65535;
    dim3 grid((unsigned)((n_points + 255) / 256), (unsigned)ng);
    one_hot_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(data + (int64_t)g0 * n_points * f, f, col0, n_classes, n_points,
                                                           n_internal, boundary_class + (int64_t)g0 * (n_points - n_internal));
    PCFD_CHECK_LAUNCH();
  }
  return PCFD_OK;
}

extern "C" int pcfd_gather_blocks(const void* src, int64_t block_bytes, const int64_t* ids, int64_t n_ids, void* dst,
                                  void* stream) {
  if (!src || !ids || !dst || block_bytes <= 0 || n_ids < 0 || block_bytes % 4) return PCFD_ERR_ARG;
  if (n_ids == 0) return PCFD_OK;
  if (n_ids > 65535) return PCFD_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool v16 = block_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
  const int64_t elems = v16 ? block_bytes / 16 : block_bytes / 4;
  // enough CTAs per block to fill the machine, each thread a few 16-byte copies in flight
  int64_t per = (elems + 256 * 4 - 1) / (256 * 4);
  const int64_t want = (148 * 8 + n_ids - 1) / n_ids;
  if (per > want) per = want;
  if (per < 1) per = 1;
  dim3 grid((unsigned)per, (unsigned)n_ids);
  if (v16) gather_blocks_kernel<uint4><<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(src), elems, ids, reinterpret_cast<uint4*>(dst));
  else gather_blocks_kernel<uint32_t><<<grid, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(src), elems, ids, reinterpret_cast<uint32_t*>(dst));
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

extern "C" int pcfd_gather_blocks_multi(const void* const* src_host, void* const* dst_host, const int64_t* block_bytes_host,
                                        int32_t n_tensors, const int64_t* ids, int64_t n_ids, void* stream) {
  if (!src_host || !dst_host || !block_bytes_host || !ids || n_tensors <= 0 || n_tensors > PCFD_GATHER_MAX_TENSORS)
    return PCFD_ERR_ARG;
  if (n_ids < 0 || n_ids > 65535) return PCFD_ERR_ARG;
  GatherTable t;
  int m = 0;
  int64_t largest = 0;
  for (int i = 0; i < n_tensors; ++i) {
    const int64_t bytes = block_bytes_host[i];
    if (bytes < 0 || bytes % 4) return PCFD_ERR_ARG;
    if (bytes == 0) continue;                                   // an empty sub-domain (no observation points)
    if (!src_host[i] || !dst_host[i]) return PCFD_ERR_ARG;
    const bool v16 = bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(src_host[i]) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(dst_host[i]) & 15) == 0;
    t.src[m] = src_host[i]; t.dst[m] = dst_host[i]; t.vec[m] = v16 ? 1 : 0;
    t.elems[m] = v16 ? bytes / 16 : bytes / 4;
    if (t.elems[m] > largest) largest = t.elems[m];
    ++m;
  }
  if (m == 0 || n_ids == 0) return PCFD_OK;
  int64_t per = (largest + 256 * 4 - 1) / (256 * 4);
  const int64_t want = (148 * 8 + n_ids - 1) / n_ids;
  if (per > want) per = want;
  if (per < 1) per = 1;
  dim3 grid((unsigned)per, (unsigned)n_ids, (unsigned)m);
  gather_blocks_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t, ids);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}
