// Jet layers whose shape makes a tensor-core tile pointless (fp32 FFMA, exact):
//
//   thin_n   n <= 8 outputs     (the last layer of every point chain: Linear(128, D+1), Linear(352, 3), ...):
//            forward = CJ*n dot products per row, 8 lanes per row; dX = a rank-n update per row followed by the
//            reverse activation jet -- both HBM-bound streams over the wide operand.
//   thin_k   k <= 16 inputs     (the first layer of every chain: Linear(D, 64), the value-only encoders' first
//            layers): forward = k FMAs per output, HBM-bound on the output.
//   small_rows  rows <= 32, cj = 1, no input transform (the per-geometry constant of the concat layer:
//            [B x 1024] x [1024 x 384]): one warp per output column / 8 contraction lanes per input column, instead
//            of a handful of 64-wide CTA tiles looping over a contraction of a thousand.
//
// The generic 64-column CTA tile of jet_linear_ffma.cu spends 16x its useful work on n = 4 and leaves 140 SMs idle
// on rows = 32 (measured: 130 us and 93 us per launch on the abc PIPN++ step; these take 5-15 us).
#include "common.cuh"

namespace pcfd {

struct ThinFwdArgs {
  const float* zin; int64_t zin_ps; int ldzin;
  InTrans tin;
  const float* w; int ldw; const float* bias; const float* cvec; int ldcvec;
  float* zout; int64_t zout_ps; int ldzout;
  int64_t rows, rows_per_geom; int k, n;
  int vec_in, vec_out;
};

struct ThinDxArgs {
  const float* gzout; int64_t gzout_ps; int ldgzout;
  const float* w; int ldw;
  const float* zin; int64_t zin_ps; int ldzin;
  InTrans tin;
  float* gzin; int64_t gzin_ps; int ldgzin;
  int64_t rows, rows_per_geom; int k, n;
  int vec_in, vec_out;
};

// ---------------------------------------------------------------------------------------------------------------
// thin_n forward: one warp per row, lanes over 4-column chunks of the contraction
// ---------------------------------------------------------------------------------------------------------------
template <int CJ>
__global__ void __launch_bounds__(256) thin_n_fwd_kernel(ThinFwdArgs a) {
  extern __shared__ float wsm[];                       // [8][k4], rows >= n and columns >= k are zero
  const int k4 = (a.k + 3) & ~3;
  for (int i = threadIdx.x; i < 8 * k4; i += 256) {
    const int j = i / k4, kk = i - j * k4;
    wsm[i] = (j < a.n && kk < a.k) ? __ldg(a.w + (int64_t)j * a.ldw + kk) : 0.0f;
  }
  __syncthreads();
  // 8 lanes per row (4 rows per warp): 3 shuffle steps per accumulator instead of 5, and exactly n <= 8 writer lanes
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & 7, rg = lane >> 3;
  const uint64_t seed = a.tin.seed_dev ? *a.tin.seed_dev : 0ULL;
  const bool plain = a.tin.act == PCFD_ACT_NONE && a.tin.escale == nullptr && a.tin.drop_p == 0.0f;
  for (int64_t row0 = ((int64_t)blockIdx.x * 8 + warp) * 4; row0 < a.rows; row0 += (int64_t)gridDim.x * 32) {
    const int64_t row = row0 + rg;
    const bool valid = row < a.rows;
    float acc[CJ][8];
#pragma unroll
    for (int c = 0; c < CJ; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[c][j] = 0.0f;
    const int64_t geom = valid ? geom_of(row, a.rows_per_geom) : 0;
    for (int col0 = sub * 4; col0 < a.k; col0 += 32) {
      float v[CJ][4];
#pragma unroll
      for (int c = 0; c < CJ; ++c) {
        const float* p = a.zin + c * a.zin_ps + row * a.ldzin + col0;
        if (!valid) {
          v[c][0] = v[c][1] = v[c][2] = v[c][3] = 0.0f;
        } else if (a.vec_in && col0 + 4 <= a.k) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(p));
          v[c][0] = x.x; v[c][1] = x.y; v[c][2] = x.z; v[c][3] = x.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[c][e] = col0 + e < a.k ? __ldg(p + e) : 0.0f;
        }
      }
      if (!plain && valid) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (col0 + e < a.tin.act_cols) {
            float zz[CJ];
#pragma unroll
            for (int c = 0; c < CJ; ++c) zz[c] = v[c][e];
            float m;
            const float s = in_scale(a.tin, seed, row, geom, col0 + e, m);
            jet_act_fwd<CJ>(a.tin.act, s, zz);
#pragma unroll
            for (int c = 0; c < CJ; ++c) v[c][e] = zz[c];
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < a.n) {
          const float4 w4 = *reinterpret_cast<const float4*>(wsm + j * k4 + col0);
#pragma unroll
          for (int c = 0; c < CJ; ++c)
            acc[c][j] = fmaf(v[c][0], w4.x, fmaf(v[c][1], w4.y, fmaf(v[c][2], w4.z, fmaf(v[c][3], w4.w, acc[c][j]))));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < a.n) {
#pragma unroll
        for (int c = 0; c < CJ; ++c) {
          float t = acc[c][j];
          t += __shfl_xor_sync(0xffffffffu, t, 1);
          t += __shfl_xor_sync(0xffffffffu, t, 2);
          t += __shfl_xor_sync(0xffffffffu, t, 4);
          acc[c][j] = t;
        }
      }
    }
    if (valid && sub < a.n) {
#pragma unroll
      for (int c = 0; c < CJ; ++c) {
        float val = acc[c][0];
#pragma unroll
        for (int j = 1; j < 8; ++j) val = sub == j ? acc[c][j] : val;
        if (c == 0) {
          if (a.bias != nullptr) val += __ldg(a.bias + sub);
          if (a.cvec != nullptr) val += __ldg(a.cvec + geom * a.ldcvec + sub);
        }
        a.zout[c * a.zout_ps + row * a.ldzout + sub] = val;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// thin_n dX: one thread per (row, 4 input columns)
// ---------------------------------------------------------------------------------------------------------------
template <int CJ>
__global__ void __launch_bounds__(256) thin_n_dx_kernel(ThinDxArgs a) {
  extern __shared__ float wsm[];                       // [8][k4]
  const int k4 = (a.k + 3) & ~3;
  for (int i = threadIdx.x; i < 8 * k4; i += 256) {
    const int j = i / k4, kk = i - j * k4;
    wsm[i] = (j < a.n && kk < a.k) ? __ldg(a.w + (int64_t)j * a.ldw + kk) : 0.0f;
  }
  __syncthreads();
  const uint64_t seed = a.tin.seed_dev ? *a.tin.seed_dev : 0ULL;
  const bool plain = a.tin.act == PCFD_ACT_NONE && a.tin.escale == nullptr && a.tin.drop_p == 0.0f;
  const uint32_t kchunks = (uint32_t)(k4 >> 2);
  const uint32_t total = (uint32_t)a.rows * kchunks;
  for (uint32_t idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    const uint32_t r32 = idx / kchunks;
    const int col0 = (int)(idx - r32 * kchunks) * 4;
    const int64_t row = r32;
    float g[CJ][4];
#pragma unroll
    for (int c = 0; c < CJ; ++c) { g[c][0] = 0.f; g[c][1] = 0.f; g[c][2] = 0.f; g[c][3] = 0.f; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < a.n) {
        const float4 w4 = *reinterpret_cast<const float4*>(wsm + j * k4 + col0);
#pragma unroll
        for (int c = 0; c < CJ; ++c) {
          const float gz = __ldg(a.gzout + c * a.gzout_ps + row * a.ldgzout + j);
          g[c][0] = fmaf(gz, w4.x, g[c][0]); g[c][1] = fmaf(gz, w4.y, g[c][1]);
          g[c][2] = fmaf(gz, w4.z, g[c][2]); g[c][3] = fmaf(gz, w4.w, g[c][3]);
        }
      }
    }
    if (!plain && col0 < a.tin.act_cols) {
      const int64_t geom = geom_of(row, a.rows_per_geom);
      float z[CJ][4];
#pragma unroll
      for (int c = 0; c < CJ; ++c) {
        const float* p = a.zin + c * a.zin_ps + row * a.ldzin + col0;
        if (a.vec_in && col0 + 4 <= a.k) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(p));
          z[c][0] = x.x; z[c][1] = x.y; z[c][2] = x.z; z[c][3] = x.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) z[c][e] = col0 + e < a.k ? __ldg(p + e) : 0.0f;
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (col0 + e < a.tin.act_cols) {
          float zz[CJ], gg[CJ];
#pragma unroll
          for (int c = 0; c < CJ; ++c) { zz[c] = z[c][e]; gg[c] = g[c][e]; }
          float m;
          const float s = in_scale(a.tin, seed, row, geom, col0 + e, m);
          jet_act_bwd<CJ>(a.tin.act, s, m, zz, gg);
#pragma unroll
          for (int c = 0; c < CJ; ++c) g[c][e] = gg[c];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CJ; ++c) {
      float* p = a.gzin + c * a.gzin_ps + row * a.ldgzin + col0;
      if (a.vec_out && col0 + 4 <= a.ldgzin) {
        *reinterpret_cast<float4*>(p) = make_float4(g[c][0], g[c][1], g[c][2], g[c][3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col0 + e < a.k) p[e] = g[c][e];
      }
    }
  }
}

// thin_n dX with the branch-scaling gradient (PI-GANO reduction layer, models/pi_gano/pi_gano.py:66-69 through the
// operators of models/modules.py:232-245): same arithmetic per (row, 4 input columns), but a thread keeps its column chunk
// and walks the rows of ONE geometry, so d loss / d escale[geometry][column] is summed in registers and leaves as one atomic
// per (thread, column) instead of one per element.  Block = kchunks column chunks x (256 / kchunks) row lanes.
template <int CJ>
__global__ void __launch_bounds__(256) thin_n_dx_ge_kernel(ThinDxArgs a, float* gescale, int ldgescale, int rows_block) {
  extern __shared__ float wsm[];                       // [8][k4]
  const int k4 = (a.k + 3) & ~3;
  for (int i = threadIdx.x; i < 8 * k4; i += 256) {
    const int j = i / k4, kk = i - j * k4;
    wsm[i] = (j < a.n && kk < a.k) ? __ldg(a.w + (int64_t)j * a.ldw + kk) : 0.0f;
  }
  __syncthreads();
  const uint64_t seed = a.tin.seed_dev ? *a.tin.seed_dev : 0ULL;
  const int kchunks = k4 >> 2;
  const int lanes = 256 / kchunks;
  const int chunk = (int)threadIdx.x % kchunks, lane = (int)threadIdx.x / kchunks;
  if (lane >= lanes) return;
  const int64_t blocks_per_geom = (a.rows_per_geom + rows_block - 1) / rows_block;
  const int64_t geom = (int64_t)blockIdx.x / blocks_per_geom;
  const int64_t rb = (int64_t)blockIdx.x - geom * blocks_per_geom;
  const int64_t r_lo = geom * a.rows_per_geom + rb * rows_block;
  int64_t r_hi = r_lo + rows_block;
  if (r_hi > (geom + 1) * a.rows_per_geom) r_hi = (geom + 1) * a.rows_per_geom;
  if (r_hi > a.rows) r_hi = a.rows;
  const int col0 = chunk * 4;
  float ge_acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t row = r_lo + lane; row < r_hi; row += lanes) {
    float g[CJ][4];
#pragma unroll
    for (int c = 0; c < CJ; ++c) { g[c][0] = 0.f; g[c][1] = 0.f; g[c][2] = 0.f; g[c][3] = 0.f; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < a.n) {
        const float4 w4 = *reinterpret_cast<const float4*>(wsm + j * k4 + col0);
#pragma unroll
        for (int c = 0; c < CJ; ++c) {
          const float gz = __ldg(a.gzout + c * a.gzout_ps + row * a.ldgzout + j);
          g[c][0] = fmaf(gz, w4.x, g[c][0]); g[c][1] = fmaf(gz, w4.y, g[c][1]);
          g[c][2] = fmaf(gz, w4.z, g[c][2]); g[c][3] = fmaf(gz, w4.w, g[c][3]);
        }
      }
    }
    if (col0 < a.tin.act_cols) {
      float z[CJ][4];
#pragma unroll
      for (int c = 0; c < CJ; ++c) {
        const float* p = a.zin + c * a.zin_ps + row * a.ldzin + col0;
        if (a.vec_in && col0 + 4 <= a.k) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(p));
          z[c][0] = x.x; z[c][1] = x.y; z[c][2] = x.z; z[c][3] = x.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) z[c][e] = col0 + e < a.k ? __ldg(p + e) : 0.0f;
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (col0 + e < a.tin.act_cols) {
          float zz[CJ], gg[CJ];
#pragma unroll
          for (int c = 0; c < CJ; ++c) { zz[c] = z[c][e]; gg[c] = g[c][e]; }
          float m;
          const float s = in_scale(a.tin, seed, row, geom, col0 + e, m);
          ge_acc[e] += jet_act_bwd<CJ>(a.tin.act, s, m, zz, gg);
#pragma unroll
          for (int c = 0; c < CJ; ++c) g[c][e] = gg[c];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CJ; ++c) {
      float* p = a.gzin + c * a.gzin_ps + row * a.ldgzin + col0;
      if (a.vec_out && col0 + 4 <= a.ldgzin) {
        *reinterpret_cast<float4*>(p) = make_float4(g[c][0], g[c][1], g[c][2], g[c][3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col0 + e < a.k) p[e] = g[c][e];
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (col0 + e < a.k && ge_acc[e] != 0.0f) atomicAdd(gescale + geom * ldgescale + col0 + e, ge_acc[e]);
}

// ---------------------------------------------------------------------------------------------------------------
// thin_k forward: one thread per (row, 4 output columns)
// ---------------------------------------------------------------------------------------------------------------
template <int CJ>
__global__ void __launch_bounds__(256) thin_k_fwd_kernel(ThinFwdArgs a) {
  extern __shared__ float wsm[];                       // [k][n4] transposed weights
  const int n4 = (a.n + 3) & ~3;
  for (int i = threadIdx.x; i < a.k * n4; i += 256) {
    const int kk = i / n4, nn = i - kk * n4;
    wsm[i] = nn < a.n ? __ldg(a.w + (int64_t)nn * a.ldw + kk) : 0.0f;
  }
  __syncthreads();
  const uint64_t seed = a.tin.seed_dev ? *a.tin.seed_dev : 0ULL;
  const bool plain = a.tin.act == PCFD_ACT_NONE && a.tin.escale == nullptr && a.tin.drop_p == 0.0f;
  const uint32_t nchunks = (uint32_t)(n4 >> 2);
  const uint32_t total = (uint32_t)a.rows * nchunks;
  for (uint32_t idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    const uint32_t r32 = idx / nchunks;
    const int col0 = (int)(idx - r32 * nchunks) * 4;
    const int64_t row = r32;
    const int64_t geom = geom_of(row, a.rows_per_geom);
    float acc[CJ][4];
#pragma unroll
    for (int c = 0; c < CJ; ++c) { acc[c][0] = 0.f; acc[c][1] = 0.f; acc[c][2] = 0.f; acc[c][3] = 0.f; }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (col0 + e < a.n) {
        if (a.bias != nullptr) acc[0][e] += __ldg(a.bias + col0 + e);
        if (a.cvec != nullptr) acc[0][e] += __ldg(a.cvec + geom * a.ldcvec + col0 + e);
      }
    }
    for (int kk = 0; kk < a.k; ++kk) {
      float zz[CJ];
#pragma unroll
      for (int c = 0; c < CJ; ++c) zz[c] = __ldg(a.zin + c * a.zin_ps + row * a.ldzin + kk);
      if (!plain && kk < a.tin.act_cols) {
        float m;
        const float s = in_scale(a.tin, seed, row, geom, kk, m);
        jet_act_fwd<CJ>(a.tin.act, s, zz);
      }
      const float4 w4 = *reinterpret_cast<const float4*>(wsm + kk * n4 + col0);
#pragma unroll
      for (int c = 0; c < CJ; ++c) {
        acc[c][0] = fmaf(zz[c], w4.x, acc[c][0]); acc[c][1] = fmaf(zz[c], w4.y, acc[c][1]);
        acc[c][2] = fmaf(zz[c], w4.z, acc[c][2]); acc[c][3] = fmaf(zz[c], w4.w, acc[c][3]);
      }
    }
#pragma unroll
    for (int c = 0; c < CJ; ++c) {
      float* p = a.zout + c * a.zout_ps + row * a.ldzout + col0;
      if (a.vec_out && col0 + 4 <= a.ldzout) {
        *reinterpret_cast<float4*>(p) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col0 + e < a.n) p[e] = acc[c][e];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// small_rows (rows <= 32, cj = 1, no input transform)
// ---------------------------------------------------------------------------------------------------------------
// forward and dX share one kernel: out[r][c] = sum_j x[r][j] * W(c, j)  with W(c, j) = w[c*ldw + j] (forward: c = output
// column, j = input) or w[j*ldw + c] (dX: c = input column, j = output).  One CTA = 8 output columns (one warp
// each) x all rows (lane = row); the contraction is tiled through shared memory in steps of 128, so there are no
// shuffles and every global load is issued by all 256 threads at once.
template <bool TRANSPOSED_W>
__global__ void __launch_bounds__(256) small_rows_kernel(const float* __restrict__ x, int ldx,
                                                         const float* __restrict__ w, int ldw, const float* bias,
                                                         const float* cvec, int ldcvec, float* out, int ldout, int rows,
                                                         int64_t rows_per_geom, int kdim, int ncols) {
  __shared__ float xs[32][129];
  __shared__ float wt[8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 8;
  float acc = 0.0f;
  for (int j0 = 0; j0 < kdim; j0 += 128) {
    for (int i = threadIdx.x; i < 32 * 128; i += 256) {
      const int r = i >> 7, jj = i & 127;
      xs[r][jj] = (r < rows && j0 + jj < kdim) ? __ldg(x + (int64_t)r * ldx + j0 + jj) : 0.0f;
    }
    for (int i = threadIdx.x; i < 8 * 128; i += 256) {
      int cc, jj;
      if (TRANSPOSED_W) { cc = i & 7; jj = i >> 3; } else { cc = i >> 7; jj = i & 127; }
      float v = 0.0f;
      if (c0 + cc < ncols && j0 + jj < kdim)
        v = TRANSPOSED_W ? __ldg(w + (int64_t)(j0 + jj) * ldw + c0 + cc) : __ldg(w + (int64_t)(c0 + cc) * ldw + j0 + jj);
      wt[cc][jj] = v;
    }
    __syncthreads();
#pragma unroll 16
    for (int jj = 0; jj < 128; ++jj) acc = fmaf(xs[lane][jj], wt[warp][jj], acc);
    __syncthreads();
  }
  const int c = c0 + warp;
  if (lane < rows && c < ncols) {
    if (bias != nullptr) acc += __ldg(bias + c);
    if (cvec != nullptr) acc += __ldg(cvec + geom_of(lane, rows_per_geom) * ldcvec + c);
    out[(int64_t)lane * ldout + c] = acc;
  }
}

// dW (+ bias / per-geometry constant gradients): one thread per weight
__global__ void __launch_bounds__(256) small_rows_dw_kernel(const float* __restrict__ gz, int ldgz,
                                                            const float* __restrict__ zin, int ldzin, float* gw, int ldgw,
                                                            float* gbias, float* gcvec, int ldgcvec, int rows,
                                                            int64_t rows_per_geom, int k, int n) {
  const uint32_t total = (uint32_t)n * (uint32_t)k;
  for (uint32_t idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    const int j = (int)(idx / (uint32_t)k), kk = (int)(idx - (uint32_t)j * (uint32_t)k);
    float s = 0.0f;
    for (int r = 0; r < rows; ++r) s = fmaf(__ldg(gz + (int64_t)r * ldgz + j), __ldg(zin + (int64_t)r * ldzin + kk), s);
    if (gw != nullptr) gw[(int64_t)j * ldgw + kk] += s;
    if (kk == 0 && (gbias != nullptr || gcvec != nullptr)) {
      float b = 0.0f;
      for (int r = 0; r < rows; ++r) {
        const float v = __ldg(gz + (int64_t)r * ldgz + j);
        b += v;
        if (gcvec != nullptr) gcvec[geom_of(r, rows_per_geom) * ldgcvec + j] += v;
      }
      if (gbias != nullptr) gbias[j] += b;
    }
  }
}

static inline int grid_for(int64_t work_items, int per_block) {
  int64_t g = (work_items + per_block - 1) / per_block;
  const int64_t cap = 148 * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace pcfd

using namespace pcfd;

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// which = 0 thin_n, 1 thin_k, 2 small_rows, -1 none
extern "C" int pcfd_thin_fwd_kind(const pcfd_intrans_t* tin, int32_t cj, int64_t rows, int32_t k, int32_t n) {
  if (!valid_cj(cj) || rows >= ((int64_t)1 << 31) / 512) return -1;
  if (rows <= 32 && cj == 1 && tin == nullptr) return 2;
  if (n <= 8 && k <= 1536) return 0;
  if (k <= 16 && n <= 512) return 1;
  return -1;
}

extern "C" int pcfd_thin_jet_linear_fwd(const float* zin, int64_t zin_ps, int32_t ldzin, const pcfd_intrans_t* tin,
                                        const float* w, int32_t ldw, const float* bias, const float* cvec,
                                        int32_t ldcvec, float* zout, int64_t zout_ps, int32_t ldzout, int32_t cj,
                                        int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int kind = pcfd_thin_fwd_kind(tin, cj, rows, k, n);
  if (kind < 0) return PCFD_ERR_ARG;
  if (kind == 2) {
    small_rows_kernel<false><<<(n + 7) / 8, 256, 0, st>>>(zin, ldzin, w, ldw, bias, cvec, ldcvec, zout, ldzout, (int)rows,
                                                          rows_per_geom, k, n);
    PCFD_CHECK_LAUNCH();
    return PCFD_OK;
  }
  ThinFwdArgs a{zin, zin_ps, ldzin, make_intrans(tin, k), w, ldw, bias, cvec, ldcvec, zout, zout_ps, ldzout,
                rows, rows_per_geom, k, n, 0, 0};
  a.vec_in = al16(zin) && ldzin % 4 == 0 && (cj == 1 || zin_ps % 4 == 0);
  a.vec_out = al16(zout) && ldzout % 4 == 0 && (cj == 1 || zout_ps % 4 == 0);
  if (kind == 0) {
    const int smem = 8 * ((k + 3) & ~3) * 4;
    const int grid = grid_for(rows, 32);
#define PCFD_THIN(CJ_) thin_n_fwd_kernel<CJ_><<<grid, 256, smem, st>>>(a); break;
    switch (cj) { case 1: PCFD_THIN(1) case 3: PCFD_THIN(3) case 4: PCFD_THIN(4) case 5: PCFD_THIN(5) case 7: PCFD_THIN(7) }
#undef PCFD_THIN
  } else {
    const int n4 = (n + 3) & ~3;
    const int smem = k * n4 * 4;
    const int grid = grid_for(rows * (n4 / 4), 256);
#define PCFD_THIN(CJ_) thin_k_fwd_kernel<CJ_><<<grid, 256, smem, st>>>(a); break;
    switch (cj) { case 1: PCFD_THIN(1) case 3: PCFD_THIN(3) case 4: PCFD_THIN(4) case 5: PCFD_THIN(5) case 7: PCFD_THIN(7) }
#undef PCFD_THIN
  }
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

// kind 3: thin_n with the branch-scaling gradient (thin_n_dx_ge_kernel): needs one thread per column chunk in a block
extern "C" int pcfd_thin_dx_kind(const pcfd_intrans_t* tin, const float* gescale, int32_t cj, int64_t rows, int32_t k,
                                 int32_t n) {
  if (!valid_cj(cj) || rows >= ((int64_t)1 << 31) / 512) return -1;
  if (gescale != nullptr) return (n <= 8 && k <= 1024 && tin != nullptr && tin->escale != nullptr) ? 3 : -1;
  if (rows <= 32 && cj == 1 && tin == nullptr) return 2;
  if (n <= 8 && k <= 1536) return 0;
  return -1;
}

extern "C" int pcfd_thin_jet_linear_bwd_dx(const float* gzout, int64_t gzout_ps, int32_t ldgzout, const float* w,
                                           int32_t ldw, const float* zin, int64_t zin_ps, int32_t ldzin,
                                           const pcfd_intrans_t* tin, float* gzin, int64_t gzin_ps, int32_t ldgzin,
                                           float* gescale, int32_t ldgescale,
                                           int32_t cj, int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n,
                                           void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int kind = pcfd_thin_dx_kind(tin, gescale, cj, rows, k, n);
  if (kind < 0) return PCFD_ERR_ARG;
  if (kind == 2) {
    small_rows_kernel<true><<<(k + 7) / 8, 256, 0, st>>>(gzout, ldgzout, w, ldw, nullptr, nullptr, 0, gzin, ldgzin,
                                                         (int)rows, 0, n, k);
    PCFD_CHECK_LAUNCH();
    return PCFD_OK;
  }
  ThinDxArgs a{gzout, gzout_ps, ldgzout, w, ldw, zin, zin_ps, ldzin, make_intrans(tin, k), gzin, gzin_ps, ldgzin,
               rows, rows_per_geom, k, n, 0, 0};
  a.vec_in = al16(zin) && ldzin % 4 == 0 && (cj == 1 || zin_ps % 4 == 0);
  a.vec_out = al16(gzin) && ldgzin % 4 == 0 && (cj == 1 || gzin_ps % 4 == 0);
  const int k4 = (k + 3) & ~3;
  const int smem = 8 * k4 * 4;
  if (kind == 3) {
    if (rows_per_geom <= 0 || rows % rows_per_geom != 0) return PCFD_ERR_ARG;
    // rows of a geometry in blocks of (row lanes x 32): 32 rows per thread amortise the atomics of the scaling gradient
    const int lanes = 256 / (k4 / 4);
    int rows_block = lanes * 32;
    const int64_t n_geom = rows / rows_per_geom;
    const int64_t blocks_per_geom = (rows_per_geom + rows_block - 1) / rows_block;
    const unsigned gridg = (unsigned)(n_geom * blocks_per_geom);
#define PCFD_THIN_GE(CJ_) thin_n_dx_ge_kernel<CJ_><<<gridg, 256, smem, st>>>(a, gescale, ldgescale, rows_block); break;
    switch (cj) { case 1: PCFD_THIN_GE(1) case 3: PCFD_THIN_GE(3) case 4: PCFD_THIN_GE(4) case 5: PCFD_THIN_GE(5) case 7: PCFD_THIN_GE(7) }
#undef PCFD_THIN_GE
    PCFD_CHECK_LAUNCH();
    return PCFD_OK;
  }
  const int grid = grid_for(rows * (k4 / 4), 256);
#define PCFD_THIN(CJ_) thin_n_dx_kernel<CJ_><<<grid, 256, smem, st>>>(a); break;
  switch (cj) { case 1: PCFD_THIN(1) case 3: PCFD_THIN(3) case 4: PCFD_THIN(4) case 5: PCFD_THIN(5) case 7: PCFD_THIN(7) }
#undef PCFD_THIN
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

extern "C" int pcfd_small_rows_supported_dw(const pcfd_intrans_t* tin, int32_t cj, int64_t rows, int32_t k, int32_t n) {
  return rows <= 32 && cj == 1 && tin == nullptr && (int64_t)k * n < ((int64_t)1 << 31);
}

extern "C" int pcfd_small_rows_bwd_dw(const float* gzout, int32_t ldgzout, const float* zin, int32_t ldzin, float* gw,
                                      int32_t ldgw, float* gbias, float* gcvec, int32_t ldgcvec, int64_t rows,
                                      int64_t rows_per_geom, int32_t k, int32_t n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (gcvec != nullptr && rows_per_geom <= 0) return PCFD_ERR_ARG;
  const int grid = grid_for((int64_t)n * k, 256);
  small_rows_dw_kernel<<<grid, 256, 0, st>>>(gzout, ldgzout, zin, ldzin, gw, ldgw, gbias, gcvec, ldgcvec, (int)rows,
                                             rows_per_geom, k, n);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}
