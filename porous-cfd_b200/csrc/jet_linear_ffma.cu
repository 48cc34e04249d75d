// Jet linear layer on CUDA cores (fp32 FFMA): forward, backward-to-input, backward-to-weights.
//
// This is the full-precision engine (engine 0): every product is an fp32 FMA, so results match
// the reference's fp32 CPU path to summation-order noise.  The tcgen05 engine (jet_linear_tc.cu)
// shares the jet algebra in common.cuh and is validated against this one.
//
// Tiling (all three kernels, 256 threads = 16 x 16):
//   * a CTA owns MP = 16*PPT points x ALL cj channels of those points, so the input transform
//     (activation jet, dropout, branch scaling) and its reverse see whole jets in one thread;
//   * shared-memory tiles are stored contraction-major ([kk][channel][point] / [kk][col]) so the
//     inner loop reads float4 vectors without bank conflicts;
//   * global loads of the next tile are issued before the FMAs of the current one.
#include "common.cuh"

namespace pcfd {

constexpr int BK = 16;   // contraction step
constexpr int BN = 64;   // output columns per CTA
constexpr int PAD = 4;

struct FwdArgs {
  const float* zin; int64_t zin_ps; int ldzin;
  InTrans tin;
  const float* w; int ldw; const float* bias; const float* cvec; int ldcvec;
  float* zout; int64_t zout_ps; int ldzout;
  int64_t rows, rows_per_geom; int k, n; int vec_out;
};

template <int CJ, int PPT>
__global__ void __launch_bounds__(256) jet_fwd_kernel(FwdArgs a) {
  constexpr int MP = 16 * PPT;
  constexpr int AS = CJ * MP + PAD;
  constexpr int BS = BN + PAD;
  __shared__ __align__(16) float As[BK * AS];
  __shared__ __align__(16) float Bs[BK * BS];

  const int tid = threadIdx.x;
  const int lk = tid & 15, lp = tid >> 4;       // loader coordinates
  const int tx = tid & 15, ty = tid >> 4;       // compute coordinates
  const int64_t row0 = (int64_t)blockIdx.x * MP;
  const int n0 = blockIdx.y * BN;
  const uint64_t seed = a.tin.seed_dev ? *a.tin.seed_dev : 0ULL;
  const bool plain = (a.tin.act == PCFD_ACT_NONE && a.tin.escale == nullptr && a.tin.drop_p == 0.0f);

  int64_t lrow[PPT]; int64_t lgeom[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    lrow[i] = row0 + lp + 16 * i;
    lgeom[i] = geom_of(lrow[i], a.rows_per_geom);
  }

  float ra[PPT][CJ];
  float rb[4];
  auto prefetch = [&](int k0) {
    const int col = k0 + lk;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const bool ok = (lrow[i] < a.rows) && (col < a.k);
#pragma unroll
      for (int c = 0; c < CJ; ++c)
        ra[i][c] = ok ? __ldg(a.zin + c * a.zin_ps + lrow[i] * a.ldzin + col) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int nn = n0 + lp + 16 * i;
      rb[i] = (nn < a.n && col < a.k) ? __ldg(a.w + (int64_t)nn * a.ldw + col) : 0.0f;
    }
  };
  auto stage = [&](int k0) {
    const int col = k0 + lk;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      if (!plain && col < a.tin.act_cols && lrow[i] < a.rows) {
        float m;
        float s = in_scale(a.tin, seed, lrow[i], lgeom[i], col, m);
        jet_act_fwd<CJ>(a.tin.act, s, ra[i]);
      }
#pragma unroll
      for (int c = 0; c < CJ; ++c) As[lk * AS + c * MP + lp + 16 * i] = ra[i][c];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[lk * BS + lp + 16 * i] = rb[i];
  };

  float acc[CJ][PPT][4];
#pragma unroll
  for (int c = 0; c < CJ; ++c)
#pragma unroll
    for (int i = 0; i < PPT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[c][i][j] = 0.0f;

  prefetch(0);
  for (int k0 = 0; k0 < a.k; k0 += BK) {
    stage(k0);
    __syncthreads();
    if (k0 + BK < a.k) prefetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk * BS + tx * 4]);
#pragma unroll
      for (int c = 0; c < CJ; ++c) {
        float av[PPT];
#pragma unroll
        for (int q = 0; q < PPT / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(&As[kk * AS + c * MP + ty * PPT + q * 4]);
          av[q * 4 + 0] = v.x; av[q * 4 + 1] = v.y; av[q * 4 + 2] = v.z; av[q * 4 + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
          acc[c][i][0] = fmaf(av[i], b.x, acc[c][i][0]);
          acc[c][i][1] = fmaf(av[i], b.y, acc[c][i][1]);
          acc[c][i][2] = fmaf(av[i], b.z, acc[c][i][2]);
          acc[c][i][3] = fmaf(av[i], b.w, acc[c][i][3]);
        }
      }
    }
    __syncthreads();
  }

  // epilogue: bias and per-geometry constant on the value channel, store pre-activations
  const int nc = n0 + tx * 4;
  float bj[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bj[j] = (a.bias != nullptr && nc + j < a.n) ? __ldg(a.bias + nc + j) : 0.0f;
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int64_t row = row0 + ty * PPT + i;
    if (row >= a.rows) continue;
    float cv[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.cvec != nullptr) {
      const int64_t g = geom_of(row, a.rows_per_geom);
#pragma unroll
      for (int j = 0; j < 4; ++j) if (nc + j < a.n) cv[j] = __ldg(a.cvec + g * a.ldcvec + nc + j);
    }
#pragma unroll
    for (int c = 0; c < CJ; ++c) {
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = acc[c][i][j] + (c == 0 ? bj[j] + cv[j] : 0.0f);
      float* dst = a.zout + c * a.zout_ps + row * a.ldzout + nc;
      if (a.vec_out && nc + 3 < a.n) {
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (nc + j < a.n) dst[j] = o[j];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
struct DxArgs {
  const float* gzout; int64_t gzout_ps; int ldgzout;
  const float* w; int ldw;
  const float* zin; int64_t zin_ps; int ldzin;
  InTrans tin;
  float* gzin; int64_t gzin_ps; int ldgzin;
  float* gescale; int ldgescale;
  int64_t rows, rows_per_geom; int k, n; int vec_out;
};

template <int CJ, int PPT>
__global__ void __launch_bounds__(256) jet_dx_kernel(DxArgs a) {
  constexpr int MP = 16 * PPT;
  constexpr int AS = CJ * MP + PAD;
  constexpr int BS = BN + PAD;
  __shared__ __align__(16) float As[BK * AS];
  __shared__ __align__(16) float Bs[BK * BS];

  const int tid = threadIdx.x;
  const int lk = tid & 15, lp = tid >> 4;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * MP;
  const int c0 = blockIdx.y * BN;           // first input column (k index) of this tile
  const uint64_t seed = a.tin.seed_dev ? *a.tin.seed_dev : 0ULL;
  const bool plain = (a.tin.act == PCFD_ACT_NONE && a.tin.escale == nullptr && a.tin.drop_p == 0.0f);

  float ra[PPT][CJ];
  float rb[4];
  auto prefetch = [&](int n0) {
    const int nn = n0 + lk;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const int64_t row = row0 + lp + 16 * i;
      const bool ok = (row < a.rows) && (nn < a.n);
#pragma unroll
      for (int c = 0; c < CJ; ++c)
        ra[i][c] = ok ? __ldg(a.gzout + c * a.gzout_ps + row * a.ldgzout + nn) : 0.0f;
    }
    // W tile: 16 rows (n) x 64 cols (k); thread (lp = n row, lk*4.. = 4 cols)
    const int wn = n0 + lp;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = c0 + lk * 4 + j;
      rb[j] = (wn < a.n && col < a.k) ? __ldg(a.w + (int64_t)wn * a.ldw + col) : 0.0f;
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int i = 0; i < PPT; ++i)
#pragma unroll
      for (int c = 0; c < CJ; ++c) As[lk * AS + c * MP + lp + 16 * i] = ra[i][c];
    *reinterpret_cast<float4*>(&Bs[lp * BS + lk * 4]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
  };

  float acc[CJ][PPT][4];
#pragma unroll
  for (int c = 0; c < CJ; ++c)
#pragma unroll
    for (int i = 0; i < PPT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[c][i][j] = 0.0f;

  prefetch(0);
  for (int n0 = 0; n0 < a.n; n0 += BK) {
    stage();
    __syncthreads();
    if (n0 + BK < a.n) prefetch(n0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk * BS + tx * 4]);
#pragma unroll
      for (int c = 0; c < CJ; ++c) {
        float av[PPT];
#pragma unroll
        for (int q = 0; q < PPT / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(&As[kk * AS + c * MP + ty * PPT + q * 4]);
          av[q * 4 + 0] = v.x; av[q * 4 + 1] = v.y; av[q * 4 + 2] = v.z; av[q * 4 + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
          acc[c][i][0] = fmaf(av[i], b.x, acc[c][i][0]);
          acc[c][i][1] = fmaf(av[i], b.y, acc[c][i][1]);
          acc[c][i][2] = fmaf(av[i], b.z, acc[c][i][2]);
          acc[c][i][3] = fmaf(av[i], b.w, acc[c][i][3]);
        }
      }
    }
    __syncthreads();
  }

  // epilogue: reverse of the input transform, then store d/dzin; accumulate d/descale
  const int kc = c0 + tx * 4;
  float ge_acc[4] = {0.f, 0.f, 0.f, 0.f};
  int64_t ge_geom = -1;
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int64_t row = row0 + ty * PPT + i;
    if (row >= a.rows) continue;
    const int64_t geom = geom_of(row, a.rows_per_geom);
    if (a.gescale != nullptr && geom != ge_geom) {
      if (ge_geom >= 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (kc + j < a.k && ge_acc[j] != 0.0f) atomicAdd(a.gescale + ge_geom * a.ldgescale + kc + j, ge_acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) ge_acc[j] = 0.0f;
      ge_geom = geom;
    }
    float out[CJ][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = kc + j;
      float g[CJ];
#pragma unroll
      for (int c = 0; c < CJ; ++c) g[c] = acc[c][i][j];
      if (col < a.k && !plain && col < a.tin.act_cols) {
        float z[CJ];
#pragma unroll
        for (int c = 0; c < CJ; ++c) z[c] = __ldg(a.zin + c * a.zin_ps + row * a.ldzin + col);
        float m;
        const float s = in_scale(a.tin, seed, row, geom, col, m);
        const float ge = jet_act_bwd<CJ>(a.tin.act, s, m, z, g);
        ge_acc[j] += ge;
      }
#pragma unroll
      for (int c = 0; c < CJ; ++c) out[c][j] = g[c];
    }
#pragma unroll
    for (int c = 0; c < CJ; ++c) {
      float* dst = a.gzin + c * a.gzin_ps + row * a.ldgzin + kc;
      if (a.vec_out && kc + 3 < a.k) {
        *reinterpret_cast<float4*>(dst) = make_float4(out[c][0], out[c][1], out[c][2], out[c][3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (kc + j < a.k) dst[j] = out[c][j];
      }
    }
  }
  if (a.gescale != nullptr && ge_geom >= 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (kc + j < a.k && ge_acc[j] != 0.0f) atomicAdd(a.gescale + ge_geom * a.ldgescale + kc + j, ge_acc[j]);
  }
}

// ---------------------------------------------------------------------------------------------
// dW: C[n][k] = sum over (channel, row) of gzout[c][row][n] * T(zin)[c][row][k], split over row
// chunks (blockIdx.y), partials reduced in fixed order by dw_reduce_kernel.
constexpr int WN = 128;  // n tile
constexpr int WK = 64;   // k tile

struct DwArgs {
  const float* gzout; int64_t gzout_ps; int ldgzout;
  const float* zin; int64_t zin_ps; int ldzin;
  InTrans tin;
  float* partial;            // [splits][n][k]
  int64_t rows, rows_per_geom, rows_per_split; int k, n; int tiles_k;
};

template <int CJ, int RP>
__global__ void __launch_bounds__(256) jet_dw_kernel(DwArgs a) {
  constexpr int E = CJ * RP;          // contraction entries per stage
  constexpr int GS = WN + PAD;
  constexpr int AS = WK + PAD;
  __shared__ __align__(16) float Gs[E * GS];
  __shared__ __align__(16) float As[E * AS];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int tile_n = blockIdx.x / a.tiles_k, tile_k = blockIdx.x % a.tiles_k;
  const int n0 = tile_n * WN, k0 = tile_k * WK;
  const int64_t r_begin = (int64_t)blockIdx.y * a.rows_per_split;
  const int64_t r_end = min(a.rows, r_begin + a.rows_per_split);
  const uint64_t seed = a.tin.seed_dev ? *a.tin.seed_dev : 0ULL;
  const bool plain = (a.tin.act == PCFD_ACT_NONE && a.tin.escale == nullptr && a.tin.drop_p == 0.0f);

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  constexpr int A_PER_THREAD = (RP * WK + 255) / 256;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += RP) {
    // gzout tile: E rows x 128 columns, plain copy
    for (int idx = tid; idx < E * WN; idx += 256) {
      const int e = idx >> 7, nn = idx & 127;
      const int c = e / RP, p = e % RP;
      const int64_t row = r0 + p;
      const int col = n0 + nn;
      Gs[e * GS + nn] = (row < r_end && col < a.n) ? __ldg(a.gzout + c * a.gzout_ps + row * a.ldgzout + col) : 0.0f;
    }
    // transformed input tile: RP points x 64 columns, all channels per thread
#pragma unroll
    for (int q = 0; q < A_PER_THREAD; ++q) {
      const int idx = tid + q * 256;
      if (idx < RP * WK) {
        const int p = idx >> 6, kk = idx & 63;
        const int64_t row = r0 + p;
        const int col = k0 + kk;
        float z[CJ];
        const bool ok = (row < r_end && col < a.k);
#pragma unroll
        for (int c = 0; c < CJ; ++c) z[c] = ok ? __ldg(a.zin + c * a.zin_ps + row * a.ldzin + col) : 0.0f;
        if (ok && !plain && col < a.tin.act_cols) {
          const int64_t geom = geom_of(row, a.rows_per_geom);
          float m;
          const float s = in_scale(a.tin, seed, row, geom, col, m);
          jet_act_fwd<CJ>(a.tin.act, s, z);
        }
#pragma unroll
        for (int c = 0; c < CJ; ++c) As[(c * RP + p) * AS + kk] = z[c];
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float4 g0 = *reinterpret_cast<const float4*>(&Gs[e * GS + ty * 8]);
      const float4 g1 = *reinterpret_cast<const float4*>(&Gs[e * GS + ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&As[e * AS + tx * 4]);
      const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = fmaf(gv[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(gv[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(gv[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(gv[i], b.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
  float* dst = a.partial + (int64_t)blockIdx.y * a.n * a.k;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int nn = n0 + ty * 8 + i;
    if (nn >= a.n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = k0 + tx * 4 + j;
      if (kk < a.k) dst[(int64_t)nn * a.k + kk] = acc[i][j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int CJ, int PPT>
static int launch_fwd(const FwdArgs& a, cudaStream_t st) {
  constexpr int MP = 16 * PPT;
  dim3 grid((unsigned)((a.rows + MP - 1) / MP), (unsigned)((a.n + BN - 1) / BN));
  jet_fwd_kernel<CJ, PPT><<<grid, 256, 0, st>>>(a);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}
template <int CJ, int PPT>
static int launch_dx(const DxArgs& a, cudaStream_t st) {
  constexpr int MP = 16 * PPT;
  dim3 grid((unsigned)((a.rows + MP - 1) / MP), (unsigned)((a.k + BN - 1) / BN));
  jet_dx_kernel<CJ, PPT><<<grid, 256, 0, st>>>(a);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

struct DwPlan { int tiles_n, tiles_k, splits; int64_t rows_per_split; int64_t chunks, rows_per_chunk; };
static DwPlan plan_dw(int cj, int64_t rows, int64_t rows_per_geom, int k, int n) {
  DwPlan p;
  p.tiles_n = (n + WN - 1) / WN;
  p.tiles_k = (k + WK - 1) / WK;
  const int rp = cj == 1 ? 16 : 4;
  const int tiles = p.tiles_n * p.tiles_k;
  int64_t splits = (2 * 148 + tiles - 1) / tiles;
  const int64_t max_splits = (rows + 64 * rp - 1) / (64 * rp);   // at least 64 stages per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t rps = (rows + splits - 1) / splits;
  rps = (rps + rp - 1) / rp * rp;
  p.rows_per_split = rps;
  p.splits = (int)((rows + rps - 1) / rps);
  if (rows_per_geom > 0) { p.rows_per_chunk = rows_per_geom; }
  else { p.rows_per_chunk = 2048; }
  p.chunks = (rows + p.rows_per_chunk - 1) / p.rows_per_chunk;
  return p;
}

}  // namespace pcfd

using namespace pcfd;

extern "C" int pcfd_ffma_jet_linear_fwd(const float* zin, int64_t zin_ps, int32_t ldzin, const pcfd_intrans_t* tin,
                                        const float* w, int32_t ldw, const float* bias, const float* cvec,
                                        int32_t ldcvec, float* zout, int64_t zout_ps, int32_t ldzout, int32_t cj,
                                        int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n, void* stream) {
  FwdArgs a{zin, zin_ps, ldzin, make_intrans(tin, k), w, ldw, bias, cvec, ldcvec, zout, zout_ps, ldzout,
            rows, rows_per_geom, k, n, 0};
  a.vec_out = aligned16(zout) && (ldzout % 4 == 0) && (zout_ps % 4 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  switch (cj) {
    case 1: return launch_fwd<1, 8>(a, st);
    case 3: return launch_fwd<3, 4>(a, st);
    case 4: return launch_fwd<4, 4>(a, st);
    case 5: return launch_fwd<5, 4>(a, st);
    case 7: return launch_fwd<7, 4>(a, st);
  }
  return PCFD_ERR_ARG;
}

extern "C" int pcfd_ffma_jet_linear_bwd_dx(const float* gzout, int64_t gzout_ps, int32_t ldgzout, const float* w,
                                           int32_t ldw, const float* zin, int64_t zin_ps, int32_t ldzin,
                                           const pcfd_intrans_t* tin, float* gzin, int64_t gzin_ps, int32_t ldgzin,
                                           float* gescale, int32_t ldgescale, int32_t cj, int64_t rows,
                                           int64_t rows_per_geom, int32_t k, int32_t n, void* stream) {
  DxArgs a{gzout, gzout_ps, ldgzout, w, ldw, zin, zin_ps, ldzin, make_intrans(tin, k), gzin, gzin_ps, ldgzin,
           gescale, ldgescale, rows, rows_per_geom, k, n, 0};
  a.vec_out = aligned16(gzin) && (ldgzin % 4 == 0) && (gzin_ps % 4 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  switch (cj) {
    case 1: return launch_dx<1, 8>(a, st);
    case 3: return launch_dx<3, 4>(a, st);
    case 4: return launch_dx<4, 4>(a, st);
    case 5: return launch_dx<5, 4>(a, st);
    case 7: return launch_dx<7, 4>(a, st);
  }
  return PCFD_ERR_ARG;
}

extern "C" int pcfd_dw_finish(const float*, int, const float*, int32_t, float*, int32_t, float*, float*, int32_t, int64_t,
                              int64_t, int32_t, int32_t, float*, const float*, int, void*);   // dw_finish.cu

extern "C" size_t pcfd_ffma_dw_workspace_bytes(int32_t cj, int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n) {
  if (!valid_cj(cj) || rows <= 0 || k <= 0 || n <= 0) return 0;
  DwPlan p = plan_dw(cj, rows, rows_per_geom, k, n);
  return ((size_t)p.splits * n * k + (size_t)p.chunks * ((p.rows_per_chunk + 127) / 128) * n) * sizeof(float) + 256;
}

extern "C" int pcfd_ffma_jet_linear_bwd_dw(const float* gzout, int64_t gzout_ps, int32_t ldgzout, const float* zin,
                                           int64_t zin_ps, int32_t ldzin, const pcfd_intrans_t* tin, float* gw,
                                           int32_t ldgw, float* gbias, float* gcvec, int32_t ldgcvec, int32_t cj,
                                           int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  if (workspace_bytes < pcfd_ffma_dw_workspace_bytes(cj, rows, rows_per_geom, k, n)) return PCFD_ERR_WORKSPACE;
  DwPlan p = plan_dw(cj, rows, rows_per_geom, k, n);
  float* partial = reinterpret_cast<float*>(workspace);
  float* tmp = partial + (size_t)p.splits * n * k;
  cudaStream_t st = (cudaStream_t)stream;
  if (gw != nullptr) {
    DwArgs a{gzout, gzout_ps, ldgzout, zin, zin_ps, ldzin, make_intrans(tin, k), partial,
             rows, rows_per_geom, p.rows_per_split, k, n, p.tiles_k};
    dim3 grid((unsigned)(p.tiles_n * p.tiles_k), (unsigned)p.splits);
    switch (cj) {
      case 1: jet_dw_kernel<1, 16><<<grid, 256, 0, st>>>(a); break;
      case 3: jet_dw_kernel<3, 4><<<grid, 256, 0, st>>>(a); break;
      case 4: jet_dw_kernel<4, 4><<<grid, 256, 0, st>>>(a); break;
      case 5: jet_dw_kernel<5, 4><<<grid, 256, 0, st>>>(a); break;
      case 7: jet_dw_kernel<7, 4><<<grid, 256, 0, st>>>(a); break;
      default: return PCFD_ERR_ARG;
    }
    PCFD_CHECK_LAUNCH();
  }
  return pcfd_dw_finish(gw != nullptr ? partial : nullptr, p.splits, gzout, ldgzout, gw, ldgw, gbias, gcvec, ldgcvec, rows,
                        rows_per_geom, k, n, tmp, nullptr, 0, stream);
}
