// Segmented max with arg-max (global max-pool of the PointNet encoders, max aggregation of the
// set-abstraction convolution) and its backward scatter.  HBM-bound: every input element is read
// once, coalesced along the feature axis; one CTA = 32 features x 8 row lanes of one segment.
#include "common.cuh"

namespace pcfd {

__global__ void __launch_bounds__(256) segmax_fwd_kernel(const float* __restrict__ z, int ldz, int act,
                                                         const int32_t* __restrict__ slots, int64_t n_seg,
                                                         int seg_len, int c, float* __restrict__ out, int ldout,
                                                         int32_t* __restrict__ arg, float* __restrict__ zsel, int ldzsel) {
  __shared__ float sv[8][33];
  __shared__ float sz[8][33];
  __shared__ int si[8][33];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int col = blockIdx.y * 32 + lane;
  const int64_t seg = blockIdx.x;
  float best = -INFINITY, bestz = 0.0f;
  int besti = -1;
  if (col < c) {
    const float* base = z + seg * (int64_t)seg_len * ldz + col;
    for (int j = wy; j < seg_len; j += 8) {
      if (slots != nullptr && __ldg(slots + seg * seg_len + j) < 0) continue;
      const float zz = __ldg(base + (int64_t)j * ldz);
      const float v = act_value(act, zz);
      if (v > best || besti < 0) { best = v; besti = j; bestz = zz; }   // strictly greater: first maximum wins inside a lane
    }
  }
  sv[wy][lane] = best;
  sz[wy][lane] = bestz;
  si[wy][lane] = besti;
  __syncthreads();
  if (wy == 0 && col < c) {
    float b = sv[0][lane], bz = sz[0][lane];
    int bi = si[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      const float v = sv[i][lane];
      const int vi = si[i][lane];
      if (vi >= 0 && (bi < 0 || v > b || (v == b && vi < bi))) { b = v; bi = vi; bz = sz[i][lane]; }
    }
    out[seg * ldout + col] = bi >= 0 ? b : 0.0f;
    arg[seg * c + col] = bi;
    if (zsel != nullptr) zsel[seg * ldzsel + col] = bi >= 0 ? bz : 0.0f;
  }
}

__global__ void __launch_bounds__(256) segmax_bwd_kernel(const float* __restrict__ gout, int ldgout,
                                                         const int32_t* __restrict__ arg,
                                                         const float* __restrict__ z, int ldz, int act,
                                                         int64_t n_seg, int seg_len, int c,
                                                         float* __restrict__ gz, int ldgz) {
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int col = blockIdx.y * 32 + lane;
  const int64_t seg = blockIdx.x;
  if (col >= c) return;
  const int a = __ldg(arg + seg * c + col);
  const float g = __ldg(gout + seg * ldgout + col);
  for (int j = wy; j < seg_len; j += 8) {
    const int64_t row = seg * (int64_t)seg_len + j;
    float v = 0.0f;
    if (j == a) v = g * act_d1(act, __ldg(z + row * ldz + col));
    gz[row * ldgz + col] = v;
  }
}

// Short segments (set-abstraction neighbourhoods: K+1 = 17 .. 65 slots): one warp per (segment, 128 features), 16-byte
// loads, the slot loop runs in registers -- no shared memory, no barrier.  Ties resolve to the lowest slot as above.
__global__ void __launch_bounds__(256) segmax_short_fwd_kernel(const float* __restrict__ z, int ldz, int act,
                                                               const int32_t* __restrict__ slots, int64_t n_seg,
                                                               int seg_len, int c, float* __restrict__ out, int ldout,
                                                               int32_t* __restrict__ arg, float* __restrict__ zsel,
                                                               int ldzsel) {
  const int lane = threadIdx.x & 31;
  const int64_t seg = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (seg >= n_seg) return;
  const int col = blockIdx.y * 128 + lane * 4;
  const bool active = col < c;            // no early exit: the whole warp takes part in the ballots below
  float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  float bz[4] = {0.f, 0.f, 0.f, 0.f};
  int bi[4] = {-1, -1, -1, -1};
  const float* base = z + seg * (int64_t)seg_len * ldz + col;
  // slot validity of up to 96 slots as three ballot words (lane j looks at slots j, j+32, j+64), so that the row loop
  // below has no dependent loads in front of its data loads and can keep several rows in flight
  unsigned valid[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu};
  if (slots != nullptr) {
#pragma unroll
    for (int w = 0; w < 3; ++w) {
      const int j = w * 32 + lane;
      const bool ok = j < seg_len && __ldg(slots + seg * seg_len + j) >= 0;
      valid[w] = __ballot_sync(0xffffffffu, ok);
    }
  }
  for (int j0 = 0; j0 < seg_len; j0 += 4) {
    float4 x[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u;
      const unsigned word = j < 32 ? valid[0] : (j < 64 ? valid[1] : valid[2]);
      ok[u] = active && j < seg_len && ((word >> (j & 31)) & 1u);
      if (ok[u]) x[u] = __ldg(reinterpret_cast<const float4*>(base + (int64_t)j * ldz));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      const int j = j0 + u;
      const float zz[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
      const float v[4] = {act_value(act, x[u].x), act_value(act, x[u].y), act_value(act, x[u].z), act_value(act, x[u].w)};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (v[e] > best[e] || bi[e] < 0) { best[e] = v[e]; bi[e] = j; bz[e] = zz[e]; }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (active && col + e < c) {
      out[seg * ldout + col + e] = bi[e] >= 0 ? best[e] : 0.0f;
      arg[seg * c + col + e] = bi[e];
      if (zsel != nullptr) zsel[seg * ldzsel + col + e] = bi[e] >= 0 ? bz[e] : 0.0f;
    }
  }
}

__global__ void __launch_bounds__(256) segmax_short_bwd_kernel(const float* __restrict__ gout, int ldgout,
                                                               const int32_t* __restrict__ arg,
                                                               const float* __restrict__ z, int ldz, int act,
                                                               int64_t n_seg, int seg_len, int c,
                                                               float* __restrict__ gz, int ldgz) {
  const int lane = threadIdx.x & 31;
  const int64_t seg = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (seg >= n_seg) return;
  const int col = blockIdx.y * 128 + lane * 4;
  if (col >= c) return;
  int a[4];
  float g[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    a[e] = col + e < c ? __ldg(arg + seg * c + col + e) : -1;
    g[e] = col + e < c ? __ldg(gout + seg * ldgout + col + e) : 0.0f;
    if (a[e] >= 0) g[e] *= act_d1(act, __ldg(z + (seg * (int64_t)seg_len + a[e]) * ldz + col + e));
  }
  for (int j = 0; j < seg_len; ++j) {
    const int64_t row = seg * (int64_t)seg_len + j;
    *reinterpret_cast<float4*>(gz + row * ldgz + col) =
        make_float4(j == a[0] ? g[0] : 0.0f, j == a[1] ? g[1] : 0.0f, j == a[2] ? g[2] : 0.0f, j == a[3] ? g[3] : 0.0f);
  }
}

}  // namespace pcfd

using namespace pcfd;

extern "C" int pcfd_segmax_fwd_z(const float* z, int32_t ldz, int32_t act, const int32_t* slots, int64_t n_seg,
                                 int32_t seg_len, int32_t c, float* out, int32_t ldout, int32_t* arg, float* zsel,
                                 int32_t ldzsel, void* stream) {
  if (!z || !out || !arg || n_seg <= 0 || seg_len <= 0 || c <= 0 || (zsel && ldzsel < c)) return PCFD_ERR_ARG;
  if (seg_len <= 96 && ldz % 4 == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0) {
    dim3 sgrid((unsigned)((n_seg + 7) / 8), (unsigned)((c + 127) / 128));
    segmax_short_fwd_kernel<<<sgrid, 256, 0, (cudaStream_t)stream>>>(z, ldz, act, slots, n_seg, seg_len, c, out, ldout, arg,
                                                                     zsel, ldzsel);
    PCFD_CHECK_LAUNCH();
    return PCFD_OK;
  }
  dim3 grid((unsigned)n_seg, (unsigned)((c + 31) / 32));
  segmax_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, ldz, act, slots, n_seg, seg_len, c, out, ldout, arg, zsel, ldzsel);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

extern "C" int pcfd_segmax_fwd(const float* z, int32_t ldz, int32_t act, const int32_t* slots, int64_t n_seg,
                               int32_t seg_len, int32_t c, float* out, int32_t ldout, int32_t* arg, void* stream) {
  return pcfd_segmax_fwd_z(z, ldz, act, slots, n_seg, seg_len, c, out, ldout, arg, nullptr, 0, stream);
}

extern "C" int pcfd_segmax_bwd(const float* gout, int32_t ldgout, const int32_t* arg, const float* z, int32_t ldz,
                               int32_t act, int64_t n_seg, int32_t seg_len, int32_t c, float* gz, int32_t ldgz,
                               void* stream) {
  if (!gout || !arg || !z || !gz || n_seg <= 0 || seg_len <= 0 || c <= 0) return PCFD_ERR_ARG;
  if (seg_len <= 96 && ldgz % 4 == 0 && (reinterpret_cast<uintptr_t>(gz) & 15) == 0 && ldgz >= ((c + 3) & ~3)) {
    dim3 sgrid((unsigned)((n_seg + 7) / 8), (unsigned)((c + 127) / 128));
    segmax_short_bwd_kernel<<<sgrid, 256, 0, (cudaStream_t)stream>>>(gout, ldgout, arg, z, ldz, act, n_seg, seg_len, c, gz,
                                                                     ldgz);
    PCFD_CHECK_LAUNCH();
    return PCFD_OK;
  }
  dim3 grid((unsigned)n_seg, (unsigned)((c + 31) / 32));
  segmax_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gout, ldgout, arg, z, ldz, act, n_seg, seg_len, c, gz, ldgz);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}
