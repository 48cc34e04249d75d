// Shared tail of every dW engine: fixed-order (deterministic) reduction of the row-split partial products into
// the weight gradient, and the column sums of the value plane of gzout that form the bias gradient and the
// per-geometry constant ("cvec") gradient.
//
//   colsum_kernel   tmp[chunk][sub][n] = sum over the rows of a sub-block (128..512) of gzout[0][row][:]
//                   (a chunk is one geometry, or 2048 rows when the layer has no per-geometry term); HBM-bound,
//                   16-byte loads, 128 columns x 8 row lanes per CTA
//   finish_kernel   one launch, two block roles:
//                     gw[n][k]   += sum_split partial[split][n][k]            (8 split lanes per element)
//                     gcvec[c][:] += sum_sub tmp[c][sub][:],  gbias += sum_c  (8 chunk lanes per column)
#include "common.cuh"

namespace pcfd {

constexpr int CS_ROWS_MAX = 512, CS_ROWS_MIN = 128;   // workspace queries reserve tmp for 128-row sub-blocks

__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ g, int ldg, int64_t rows,
                                                     int64_t rows_per_chunk, int subs, int cs_rows, int n, float* tmp, int vec) {
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int64_t chunk = blockIdx.x / subs, sub = blockIdx.x % subs;
  const int64_t c_begin = chunk * rows_per_chunk;
  const int64_t c_end = min(rows, c_begin + rows_per_chunk);
  const int64_t r_begin = c_begin + sub * cs_rows;
  const int64_t r_end = min(c_end, r_begin + cs_rows);
  const int col = blockIdx.y * 128 + lane * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (vec && col + 4 <= n) {
    int64_t r = r_begin + wy;
    for (; r + 24 < r_end; r += 32) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(g + r * ldg + col));
      const float4 b = __ldg(reinterpret_cast<const float4*>(g + (r + 8) * ldg + col));
      const float4 c = __ldg(reinterpret_cast<const float4*>(g + (r + 16) * ldg + col));
      const float4 d = __ldg(reinterpret_cast<const float4*>(g + (r + 24) * ldg + col));
      s.x += (a.x + b.x) + (c.x + d.x); s.y += (a.y + b.y) + (c.y + d.y);
      s.z += (a.z + b.z) + (c.z + d.z); s.w += (a.w + b.w) + (c.w + d.w);
    }
    for (; r < r_end; r += 8) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(g + r * ldg + col));
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
  } else if (col < n) {
    for (int64_t r = r_begin + wy; r < r_end; r += 8) {
      const float* p = g + r * ldg + col;
      s.x += __ldg(p);
      if (col + 1 < n) s.y += __ldg(p + 1);
      if (col + 2 < n) s.z += __ldg(p + 2);
      if (col + 3 < n) s.w += __ldg(p + 3);
    }
  }
  red[wy][lane] = s;
  __syncthreads();
  if (wy == 0 && col < n) {
    float4 t = red[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) { const float4 v = red[i][lane]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
    float* o = tmp + (int64_t)blockIdx.x * n + col;
    o[0] = t.x;
    if (col + 1 < n) o[1] = t.y;
    if (col + 2 < n) o[2] = t.z;
    if (col + 3 < n) o[3] = t.w;
  }
}

__global__ void __launch_bounds__(256) dw_finish_kernel(const float* __restrict__ partial, int splits, int n, int k,
                                                        float* gw, int ldgw, int gw_blocks,
                                                        const float* __restrict__ tmp, int64_t chunks, int subs,
                                                        float* gbias, float* gcvec, int ldgcvec) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  griddep_wait();            // launched through launch_pdl right behind the dW kernel that writes `partial`
  if ((int)blockIdx.x < gw_blocks) {
    // ---- weight gradient: element = blockIdx.x * 32 + lane, split lane wy
    const int64_t total = (int64_t)n * k;
    const int64_t idx = (int64_t)blockIdx.x * 32 + lane;
    float s = 0.0f;
    if (idx < total) {
      // five independent loads in flight per thread: the reduction is latency-bound (splits / 8 values per thread)
      for (int base = wy; base < splits; base += 40) {
        float v[5];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          const int sp = base + 8 * u;
          v[u] = sp < splits ? __ldg(partial + (int64_t)sp * total + idx) : 0.0f;
        }
        s += ((v[0] + v[1]) + (v[2] + v[3])) + v[4];
      }
    }
    red[wy][lane] = s;
    __syncthreads();
    if (wy == 0 && idx < total) {
      float t = red[0][lane];
#pragma unroll
      for (int i = 1; i < 8; ++i) t += red[i][lane];
      const int nn = (int)(idx / k), kk = (int)(idx - (int64_t)nn * k);
      gw[(int64_t)nn * ldgw + kk] += t;
    }
    return;
  }
  // ---- column sums: column = (blockIdx.x - gw_blocks) * 32 + lane, chunk lane wy
  const int col = ((int)blockIdx.x - gw_blocks) * 32 + lane;
  float s = 0.0f;
  if (col < n && subs == 1 && gcvec == nullptr) {
    for (int64_t base = wy; base < chunks; base += 40) {
      float v[5];
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int64_t c = base + 8 * u;
        v[u] = c < chunks ? __ldg(tmp + c * n + col) : 0.0f;
      }
      s += ((v[0] + v[1]) + (v[2] + v[3])) + v[4];
    }
  } else if (col < n) {
    for (int64_t c = wy; c < chunks; c += 8) {
      const float* p = tmp + c * subs * n + col;
      float v = 0.0f;
      for (int u = 0; u < subs; ++u) v += p[(int64_t)u * n];
      s += v;
      if (gcvec != nullptr) gcvec[c * ldgcvec + col] += v;
    }
  }
  red[wy][lane] = s;
  __syncthreads();
  if (wy == 0 && col < n && gbias != nullptr) {
    float t = red[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) t += red[i][lane];
    gbias[col] += t;
  }
}

// Adam on flat fp32 buffers (torch.optim.Adam semantics, no weight decay / amsgrad): one elementwise pass instead of
// torch's multi-tensor kernel, which cuts a single 861k-element tensor into 14 chunks (39 us vs ~5 us).  `step` and `lr`
// live on the device so the launch is graph-capturable and the scheduler can change the rate between replays.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, const int64_t* step, const float* lr, float b1,
                                                   float b2, float eps, float grad_scale, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float t = (float)(*step);
  const float bc1 = 1.0f - powf(b1, t), bc2 = 1.0f - powf(b2, t);
  const float step_size = *lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (i + e < n) {
      const float gr = g[i + e] * grad_scale;
      const float mm = b1 * m[i + e] + (1.0f - b1) * gr;
      const float vv = b2 * v[i + e] + (1.0f - b2) * gr * gr;
      m[i + e] = mm;
      v[i + e] = vv;
      p[i + e] -= step_size * mm / (sqrtf(vv) * inv_sqrt_bc2 + eps);
    }
  }
}

__global__ void adam_advance_kernel(int64_t* step) { *step += 1; }

}  // namespace pcfd

using namespace pcfd;

extern "C" int pcfd_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t* step,
                              const float* lr, float beta1, float beta2, float eps, float grad_scale, int64_t n,
                              void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step || !lr || n <= 0) return PCFD_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  adam_advance_kernel<<<1, 1, 0, st>>>(step);
  PCFD_CHECK_LAUNCH();
  adam_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps,
                                                            grad_scale, n);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

// tmp must hold chunks * ceil(rows_per_chunk / 128) * n floats (every engine's workspace query reserves
// chunks * ceil(rows_per_chunk / 128) * n)
// `colsum_partial` ([colsum_parts][n], optional): column sums of plane 0 of gzout already formed by the dW kernel
// (engine 2); used for the bias gradient when no per-geometry gradient is wanted, instead of a pass over gzout.
extern "C" int pcfd_dw_finish(const float* partial, int splits, const float* gzout, int32_t ldgzout, float* gw,
                              int32_t ldgw, float* gbias, float* gcvec, int32_t ldgcvec, int64_t rows,
                              int64_t rows_per_geom, int32_t k, int32_t n, float* tmp, const float* colsum_partial,
                              int colsum_parts, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const bool do_gw = gw != nullptr && partial != nullptr;
  const bool do_cs = gbias != nullptr || gcvec != nullptr;
  int64_t chunks = 0;
  int subs = 0;
  const float* sums = tmp;
  if (do_cs && gcvec == nullptr && colsum_partial != nullptr) {
    sums = colsum_partial; chunks = colsum_parts; subs = 1;
  } else if (do_cs) {
    if (gcvec != nullptr && rows_per_geom <= 0) return PCFD_ERR_ARG;
    const int64_t rows_per_chunk = rows_per_geom > 0 ? rows_per_geom : 2048;
    chunks = (rows + rows_per_chunk - 1) / rows_per_chunk;
    // long sub-blocks amortise the block reduction, but small problems need enough CTAs to hide the load latency
    int cs_rows = CS_ROWS_MAX;
    while (cs_rows > CS_ROWS_MIN && chunks * ((rows_per_chunk + cs_rows - 1) / cs_rows) * ((n + 127) / 128) < 4 * 148) cs_rows >>= 1;
    subs = (int)((rows_per_chunk + cs_rows - 1) / cs_rows);
    const int vec = (reinterpret_cast<uintptr_t>(gzout) & 15) == 0 && ldgzout % 4 == 0;
    dim3 grid((unsigned)(chunks * subs), (unsigned)((n + 127) / 128));
    colsum_kernel<<<grid, 256, 0, st>>>(gzout, ldgzout, rows, rows_per_chunk, subs, cs_rows, n, tmp, vec);
    PCFD_CHECK_LAUNCH();
  }
  if (do_gw || do_cs) {
    const int gw_blocks = do_gw ? (int)(((int64_t)n * k + 31) / 32) : 0;
    const int cs_blocks = do_cs ? (n + 31) / 32 : 0;
    const cudaError_t le = launch_pdl(dw_finish_kernel, dim3((unsigned)(gw_blocks + cs_blocks)), dim3(256), (size_t)0, st,
                                      partial, splits, n, k, gw, ldgw, gw_blocks, sums, chunks, subs, gbias, gcvec, ldgcvec);
    if (le != cudaSuccess) return PCFD_ERR_CUDA + (int)le;
  }
  return PCFD_OK;
}
