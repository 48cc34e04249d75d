// Fused Navier-Stokes-Darcy residual assembly, losses and their gradient wrt the model output jet.
//
// One pass over the output jets replaces the reference's ~40 small launches and four gathers
// (models/losses.py:149-319 + models/model_base.py:191-212): every internal point computes the
// continuity and momentum residuals from (U, p, jac, lap, grad p), adds its squares to a block
// reduction and immediately writes d(sum of weighted losses)/d(jet); boundary and observation
// points do the same for the MSE terms.  HBM-bound: algorithmic bytes per internal point are
// cj*(D+1)*4 read + cj*(D+1)*4 written + the data row.
#include "common.cuh"

namespace pcfd {

constexpr int NSLOT = 8;

struct ResArgs {
  const float* data; int64_t n_rows; int f;
  const int64_t* internal_ids; int64_t ni;
  const int64_t* boundary_ids; int64_t nb;
  const int64_t* obs_ids; int64_t no;
  const float* y_int; int64_t ps; const float* y_bnd; int ldy;
  float* gy_int; float* gy_bnd;
  pcfd_residual_params_t p;
  float inv_int, inv_bnd, inv_obs;   // 1 / (n_geom * count)
  int n_geom;
  float* partial;                     // this kernel's [blocks][NSLOT]
  const float* wdev;                  // optional device-resident loss weights (ReLoBRaLo); else p.weights
  float* fields;                      // optional per-point residual map [n_geom*ni][D+1] = (momentum xD, div)
  const float* visc_extra;            // optional [n_geom*ni][D]: added to the Laplacian row sums (vanilla-PIPN coupling)
  float* gvisc;                       // optional [n_geom*ni][D]: d loss / d visc_extra
  int vec_rows;                       // jet rows are 16-byte aligned and 4 floats apart or more: float4 loads / stores
};

#define PCFD_WT(i) (a.wdev != nullptr ? __ldg(a.wdev + (i)) : P.weights[i])

// sums of the block -> dst[0..NSLOT)
__device__ __forceinline__ void block_reduce_store(float (&v)[NSLOT], float* dst) {
  __shared__ float red[8][NSLOT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) {
    float x = v[s];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[warp][s] = x;
  }
  __syncthreads();
  if (threadIdx.x < NSLOT) {
    float t = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w][threadIdx.x];
    dst[threadIdx.x] = t;
  }
}

// one output-jet row (D + 1 values, ld floats apart rows): a single 16-byte load / store when the layout allows
template <int D>
__device__ __forceinline__ void load_row(const float* p, bool vec, float (&v)[D + 1]) {
  if (vec) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = q.x; v[1] = q.y; v[2] = q.z;
    if (D == 3) v[D] = q.w;
  } else {
#pragma unroll
    for (int o = 0; o <= D; ++o) v[o] = __ldg(p + o);
  }
}
template <int D>
__device__ __forceinline__ void store_row(float* p, bool vec, const float (&v)[D + 1]) {
  if (vec) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], D == 3 ? v[D] : 0.0f);
  } else {
#pragma unroll
    for (int o = 0; o <= D; ++o) p[o] = v[o];
  }
}

// slots: [0] continuity, [1..D] momentum, [D+1..2D] |U error|, [2D+1] |p error|
// ATOMIC0: the value plane of gy is accumulated with atomics onto a zeroed buffer (the fused step kernel, where the
// observation blocks add to the same entries in no particular order); otherwise plain stores.
template <int D, int LAP, bool ATOMIC0>
__device__ __forceinline__ void internal_body(const ResArgs& a, int64_t blk, float (&sums)[NSLOT]) {
  constexpr int CJ = LAP == PCFD_LAP_TRUE ? 1 + 2 * D : 1 + D;
  const int64_t t = blk * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)a.n_geom * a.ni;
  const bool vec = a.vec_rows != 0;
  if (t < total) {
    const pcfd_residual_params_t& P = a.p;
    const int64_t g = t / a.ni, i = t % a.ni;
    const float* drow = a.data + (g * a.n_rows + a.internal_ids[t]) * a.f;   // batch['internal'] row
    const float* erow = a.data + (g * a.n_rows + i) * a.f;                   // calculate_errors pairs row i with row i
    float y[CJ][D + 1];
#pragma unroll
    for (int c = 0; c < CJ; ++c) load_row<D>(a.y_int + c * a.ps + t * a.ldy, vec, y[c]);

    const bool manu = P.loss_kind == PCFD_LOSS_MANUFACTURED;
    float su[D], mu[D], sx[D], dcoef[D], fcoef[D];
    float sp = manu ? 1.0f : P.p_std;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      su[d] = manu ? 1.0f : P.u_std[d];
      mu[d] = manu ? 0.0f : P.u_mean[d];
      sx[d] = manu ? 1.0f : P.c_std[d];
      if (P.loss_kind == PCFD_LOSS_VARIABLE) {
        dcoef[d] = P.d_min[d] + P.d_range[d] * __ldg(drow + P.col_d[d]);
        fcoef[d] = P.f_min[d] + P.f_range[d] * __ldg(drow + P.col_f[d]);
      } else {
        dcoef[d] = P.d; fcoef[d] = P.f;
      }
    }
    const float zone = __ldg(drow + P.col_zone);
    float ur[D];
    float nrm2 = 0.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) { ur[d] = su[d] * y[0][d] + mu[d]; nrm2 += ur[d] * ur[d]; }
    const float nrm = sqrtf(nrm2);

    // continuity: div = sum_i jac[i][i] * su_i / sx_i
    float div = 0.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) div += y[1 + d][d] * su[d] / sx[d];
    // laplacian row sums.  LAP_TRUE: lap[i][j] = d2 U_i / dx_j2.  LAP_REFERENCE reproduces the
    // reference call as written (models/model_base.py:195): lap[n][i][j] = delta(n,i) * jac[n][j][j]
    // for the first D points of every geometry, zero elsewhere.
    float res[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
      float conv = 0.0f;
#pragma unroll
      for (int j = 0; j < D; ++j) conv += y[1 + j][c] * (ur[j] / sx[j]);
      conv *= su[c];
      float visc = 0.0f;
      if (LAP == PCFD_LAP_TRUE) {
#pragma unroll
        for (int j = 0; j < D; ++j) visc += y[1 + D + j < CJ ? 1 + D + j : 0][c] * (1.0f / (sx[j] * sx[j]));
      } else if (i == c) {
#pragma unroll
        for (int j = 0; j < D; ++j) visc += y[1 + j][j] * (1.0f / (sx[j] * sx[j]));
      }
      if (a.visc_extra != nullptr) visc += __ldg(a.visc_extra + t * D + c);
      visc *= P.nu * su[c];
      const float pres = (sp / sx[c]) * y[1 + c][D];
      const float source = ur[c] * (dcoef[c] * P.nu + 0.5f * nrm * fcoef[c]);
      float r = conv - visc + pres + source * zone;
      if (manu) r -= __ldg(drow + P.col_f[c]);
      res[c] = r;
    }

    sums[0] = div * div;
#pragma unroll
    for (int c = 0; c < D; ++c) sums[1 + c] = res[c] * res[c];
    if (a.fields != nullptr) {          // predict_step(verbose): residuals = cat([momentum_error, div]) (models/model_base.py:250)
#pragma unroll
      for (int c = 0; c < D; ++c) a.fields[t * (D + 1) + c] = res[c];
      a.fields[t * (D + 1) + D] = div;
    }
    // MAE log values on de-standardised fields
#pragma unroll
    for (int d = 0; d < D; ++d) sums[1 + D + d] = fabsf(su[d] * (y[0][d] - __ldg(erow + P.col_u[d])));
    sums[1 + 2 * D] = fabsf(sp * (y[0][D] - __ldg(erow + P.col_p)));

    // gradient of  w0*mean(div^2) + sum_c w_{1+c}*mean(res_c^2)
    float gy[CJ][D + 1];
#pragma unroll
    for (int c = 0; c < CJ; ++c)
#pragma unroll
      for (int o = 0; o <= D; ++o) gy[c][o] = 0.0f;
    const float gdiv = 2.0f * PCFD_WT(0) * div * a.inv_int;
    float gr[D];
#pragma unroll
    for (int c = 0; c < D; ++c) gr[c] = 2.0f * PCFD_WT(1 + c) * res[c] * a.inv_int;
#pragma unroll
    for (int d = 0; d < D; ++d) gy[1 + d][d] += gdiv * su[d] / sx[d];
    float gur[D];   // gradient wrt u_raw
#pragma unroll
    for (int j = 0; j < D; ++j) gur[j] = 0.0f;
#pragma unroll
    for (int c = 0; c < D; ++c) {
#pragma unroll
      for (int j = 0; j < D; ++j) {
        gy[1 + j][c] += gr[c] * su[c] * (ur[j] / sx[j]);          // d conv_c / d jac[c][j]
        gur[j] += gr[c] * su[c] * y[1 + j][c] / sx[j];            // d conv_c / d u_raw_j
      }
      const float gv = -gr[c] * P.nu * su[c];
      if (a.gvisc != nullptr) a.gvisc[t * D + c] = gv;
      if (LAP == PCFD_LAP_TRUE) {
#pragma unroll
        for (int j = 0; j < D; ++j) gy[1 + D + j < CJ ? 1 + D + j : 0][c] += gv * (1.0f / (sx[j] * sx[j]));
      } else if (i == c) {
#pragma unroll
        for (int j = 0; j < D; ++j) gy[1 + j][j] += gv * (1.0f / (sx[j] * sx[j]));
      }
      gy[1 + c][D] += gr[c] * (sp / sx[c]);
      // source_c = ur_c * (d_c nu + 0.5 |ur| f_c)
      const float gs = gr[c] * zone;
      gur[c] += gs * (dcoef[c] * P.nu + 0.5f * nrm * fcoef[c]);
      if (nrm > 0.0f) {
        const float k = gs * ur[c] * 0.5f * fcoef[c] / nrm;
#pragma unroll
        for (int j = 0; j < D; ++j) gur[j] += k * ur[j];
      }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) gy[0][d] += gur[d] * su[d];
    if (a.gy_int != nullptr) {
      if (ATOMIC0) {
#pragma unroll
        for (int o = 0; o <= D; ++o) atomicAdd(a.gy_int + t * a.ldy + o, gy[0][o]);
      } else {
        store_row<D>(a.gy_int + t * a.ldy, vec, gy[0]);
      }
#pragma unroll
      for (int c = 1; c < CJ; ++c) store_row<D>(a.gy_int + c * a.ps + t * a.ldy, vec, gy[c]);
    }
  }
}

template <int D, int LAP>
__global__ void __launch_bounds__(256) residual_internal_kernel(ResArgs a) {
  float sums[NSLOT];
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) sums[s] = 0.0f;
  internal_body<D, LAP, false>(a, blockIdx.x, sums);
  if (a.partial != nullptr) block_reduce_store(sums, a.partial + (int64_t)blockIdx.x * NSLOT);
}

// slots: [0..D-1] (U - target)^2, [D] (p - target)^2, [D+1..2D] |U error|, [2D+1] |p error|
template <int D, bool ATOMIC0>
__device__ __forceinline__ void boundary_body(const ResArgs& a, int64_t blk, float (&sums)[NSLOT]) {
  const int64_t t = blk * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)a.n_geom * a.nb;
  if (t < total) {
    const pcfd_residual_params_t& P = a.p;
    const int64_t g = t / a.nb, j = t % a.nb;
    const float* drow = a.data + (g * a.n_rows + a.boundary_ids[t]) * a.f;
    const float* erow = a.data + (g * a.n_rows + a.ni + j) * a.f;
    const bool manu = P.loss_kind == PCFD_LOSS_MANUFACTURED;
#pragma unroll
    for (int o = 0; o <= D; ++o) {
      const float yv = __ldg(a.y_bnd + t * a.ldy + o);
      const int col = o < D ? P.col_u[o] : P.col_p;
      const float diff = yv - __ldg(drow + col);
      sums[o] = diff * diff;
      const float w = PCFD_WT(1 + D + o);
      if (ATOMIC0) atomicAdd(a.gy_bnd + t * a.ldy + o, 2.0f * w * diff * a.inv_bnd);
      else a.gy_bnd[t * a.ldy + o] = 2.0f * w * diff * a.inv_bnd;
      const float sc = manu ? 1.0f : (o < D ? P.u_std[o] : P.p_std);
      sums[D + 1 + o] = fabsf(sc * (yv - __ldg(erow + col)));
    }
  }
}

template <int D>
__global__ void __launch_bounds__(256) residual_boundary_kernel(ResArgs a) {
  float sums[NSLOT];
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) sums[s] = 0.0f;
  boundary_body<D, false>(a, blockIdx.x, sums);
  block_reduce_store(sums, a.partial + (int64_t)blockIdx.x * NSLOT);
}

// slots: [0..D-1] obs U, [D] obs p.  Adds its gradient on top of what the two kernels above wrote.
template <int D>
__device__ __forceinline__ void obs_body(const ResArgs& a, int64_t blk, float (&sums)[NSLOT]) {
  const int64_t t = blk * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)a.n_geom * a.no;
  if (t < total) {
    const pcfd_residual_params_t& P = a.p;
    const int64_t g = t / a.no;
    const int64_t id = a.obs_ids[t];                 // row in prediction order AND in data
    const float* drow = a.data + (g * a.n_rows + id) * a.f;
    const bool internal = id < a.ni;
    const float* yrow = internal ? a.y_int + (g * a.ni + id) * a.ldy : a.y_bnd + (g * a.nb + (id - a.ni)) * a.ldy;
    float* grow = internal ? a.gy_int + (g * a.ni + id) * a.ldy : a.gy_bnd + (g * a.nb + (id - a.ni)) * a.ldy;
#pragma unroll
    for (int o = 0; o <= D; ++o) {
      const int col = o < D ? P.col_u[o] : P.col_p;
      const float diff = __ldg(yrow + o) - __ldg(drow + col);
      sums[o] = diff * diff;
      atomicAdd(grow + o, 2.0f * PCFD_WT(2 + 2 * D + o) * diff * a.inv_obs);
    }
  }
}

template <int D>
__global__ void __launch_bounds__(256) residual_obs_kernel(ResArgs a) {
  float sums[NSLOT];
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) sums[s] = 0.0f;
  obs_body<D>(a, blockIdx.x, sums);
  block_reduce_store(sums, a.partial + (int64_t)blockIdx.x * NSLOT);
}

struct FinishArgs {
  const float* p_int; int b_int; const float* p_bnd; int b_bnd; const float* p_obs; int b_obs;
  int dims, data_loss; float inv_int, inv_bnd, inv_obs, inv_all; float weights[16]; const float* wdev; float* out;
};

__device__ __forceinline__ void finish_body(const FinishArgs& a) {
  __shared__ double tot[3 * NSLOT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // one warp per (region, slot) sum, three rounds of eight: fixed order -> deterministic
  for (int q = warp; q < 3 * NSLOT; q += 8) {
    const int region = q / NSLOT, slot = q % NSLOT;
    const float* p = region == 0 ? a.p_int : (region == 1 ? a.p_bnd : a.p_obs);
    const int nb = region == 0 ? a.b_int : (region == 1 ? a.b_bnd : a.b_obs);
    double s = 0.0;
    for (int b = lane; b < nb; b += 32) s += (double)__ldcg(p + (int64_t)b * NSLOT + slot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) tot[q] = s;
  }
  __syncthreads();
  if (tid == 0) {
    const int D = a.dims;
    float* out = a.out;
    for (int i = 0; i < PCFD_LOSS_OUT_FLOATS; ++i) out[i] = 0.0f;
    int n = 0;
    out[n++] = (float)(tot[0] * a.inv_int);
    for (int c = 0; c < D; ++c) out[n++] = (float)(tot[1 + c] * a.inv_int);
    for (int c = 0; c <= D; ++c) out[n++] = (float)(tot[NSLOT + c] * a.inv_bnd);
    if (a.data_loss)
      for (int c = 0; c <= D; ++c) out[n++] = (float)(tot[2 * NSLOT + c] * a.inv_obs);
    float total = 0.0f;
    for (int i = 0; i < n; ++i) { out[16 + i] = out[i] * (a.wdev != nullptr ? a.wdev[i] : a.weights[i]); total += out[16 + i]; }
    out[32] = total;
    for (int c = 0; c < D; ++c) out[33 + c] = (float)((tot[1 + D + c] + tot[NSLOT + D + 1 + c]) * a.inv_all);
    out[36] = (float)((tot[1 + 2 * D] + tot[NSLOT + 2 * D + 1]) * a.inv_all);
    out[37] = (float)n;
  }
}

__global__ void __launch_bounds__(256) residual_finish_kernel(FinishArgs a) { finish_body(a); }

// The whole residual stage of a training step in ONE launch: blocks [0, b_int) take the internal points, the next b_bnd
// the boundary points, the rest the observation points; every block leaves its partial sums in the workspace and takes a
// ticket, and the block that draws the last ticket reduces all partials in a fixed order (deterministic) into the loss
// vector and hands the ticket counter back at zero.  The value plane of gy is accumulated with atomics (observation
// points add to entries the other two roles also write), so the caller zeroes it first (one memset node).
template <int D, int LAP>
__global__ void __launch_bounds__(256) residual_step_kernel(ResArgs a, FinishArgs fa, int b_int, int b_bnd, int* ticket) {
  __shared__ int is_last;
  float sums[NSLOT];
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) sums[s] = 0.0f;
  const int blk = blockIdx.x;
  float* dst;
  if (blk < b_int) {
    internal_body<D, LAP, true>(a, blk, sums);
    dst = const_cast<float*>(fa.p_int) + (int64_t)blk * NSLOT;
  } else if (blk < b_int + b_bnd) {
    boundary_body<D, true>(a, blk - b_int, sums);
    dst = const_cast<float*>(fa.p_bnd) + (int64_t)(blk - b_int) * NSLOT;
  } else {
    obs_body<D>(a, blk - b_int - b_bnd, sums);
    dst = const_cast<float*>(fa.p_obs) + (int64_t)(blk - b_int - b_bnd) * NSLOT;
  }
  block_reduce_store(sums, dst);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (is_last) {
    __threadfence();
    finish_body(fa);
    if (threadIdx.x == 0) *ticket = 0;
  }
}

static inline int blocks_for(int64_t n) { return (int)((n + 255) / 256); }
static inline int rows_vectorisable(const float* y, const float* gy, int64_t ps, int ldy) {
  return ldy % 4 == 0 && ps % 4 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(gy) & 15) == 0;
}

}  // namespace pcfd

using namespace pcfd;

extern "C" size_t pcfd_residual_workspace_bytes(int32_t n_geom, int64_t ni, int64_t nb, int64_t no) {
  const size_t blocks = (size_t)blocks_for(n_geom * ni) + blocks_for(n_geom * nb) + blocks_for(n_geom * no) + 3;
  return blocks * NSLOT * sizeof(float);
}

extern "C" int pcfd_residual_loss_w(const float* data, int32_t n_geom, int64_t n_rows, int32_t f,
                                    const int64_t* internal_ids, int64_t ni, const int64_t* boundary_ids, int64_t nb,
                                    const int64_t* obs_ids, int64_t no, const float* y_int, int64_t y_plane_stride,
                                    const float* y_bnd, int32_t ldy, const pcfd_residual_params_t* prm,
                                    const float* weights_dev, const float* visc_extra, float* gvisc, float* gy_int,
                                    float* gy_bnd, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!data || !internal_ids || !boundary_ids || !y_int || !y_bnd || !prm || !gy_int || !gy_bnd || !out || !workspace)
    return PCFD_ERR_ARG;
  if (n_geom <= 0 || ni <= 0 || nb <= 0 || (prm->dims != 2 && prm->dims != 3) || ldy < prm->dims + 1) return PCFD_ERR_ARG;
  if (workspace_bytes < pcfd_residual_workspace_bytes(n_geom, ni, nb, no)) return PCFD_ERR_WORKSPACE;
  const bool data_loss = prm->enable_data_loss && no > 0;
  if (data_loss && !obs_ids) return PCFD_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  ResArgs a;
  a.data = data; a.n_rows = n_rows; a.f = f;
  a.internal_ids = internal_ids; a.ni = ni; a.boundary_ids = boundary_ids; a.nb = nb; a.obs_ids = obs_ids; a.no = no;
  a.y_int = y_int; a.ps = y_plane_stride; a.y_bnd = y_bnd; a.ldy = ldy; a.gy_int = gy_int; a.gy_bnd = gy_bnd;
  a.p = *prm; a.n_geom = n_geom; a.wdev = weights_dev; a.fields = nullptr; a.visc_extra = visc_extra; a.gvisc = gvisc;
  a.vec_rows = rows_vectorisable(y_int, gy_int, y_plane_stride, ldy);
  a.inv_int = 1.0f / (float)((double)n_geom * ni);
  a.inv_bnd = 1.0f / (float)((double)n_geom * nb);
  a.inv_obs = no > 0 ? 1.0f / (float)((double)n_geom * no) : 0.0f;
  float* ws = reinterpret_cast<float*>(workspace);
  const int b_int = blocks_for((int64_t)n_geom * ni), b_bnd = blocks_for((int64_t)n_geom * nb);
  const int b_obs = data_loss ? blocks_for((int64_t)n_geom * no) : 0;
  float* p_int = ws; float* p_bnd = p_int + (size_t)b_int * NSLOT; float* p_obs = p_bnd + (size_t)b_bnd * NSLOT;

  a.partial = p_int;
  const int D = prm->dims;
  if (D == 2 && prm->lap_mode == PCFD_LAP_REFERENCE) residual_internal_kernel<2, PCFD_LAP_REFERENCE><<<b_int, 256, 0, st>>>(a);
  else if (D == 2) residual_internal_kernel<2, PCFD_LAP_TRUE><<<b_int, 256, 0, st>>>(a);
  else if (prm->lap_mode == PCFD_LAP_REFERENCE) residual_internal_kernel<3, PCFD_LAP_REFERENCE><<<b_int, 256, 0, st>>>(a);
  else residual_internal_kernel<3, PCFD_LAP_TRUE><<<b_int, 256, 0, st>>>(a);
  PCFD_CHECK_LAUNCH();
  a.partial = p_bnd;
  if (D == 2) residual_boundary_kernel<2><<<b_bnd, 256, 0, st>>>(a);
  else residual_boundary_kernel<3><<<b_bnd, 256, 0, st>>>(a);
  PCFD_CHECK_LAUNCH();
  if (data_loss) {
    a.partial = p_obs;
    if (D == 2) residual_obs_kernel<2><<<b_obs, 256, 0, st>>>(a);
    else residual_obs_kernel<3><<<b_obs, 256, 0, st>>>(a);
    PCFD_CHECK_LAUNCH();
  }
  FinishArgs fa;
  fa.p_int = p_int; fa.b_int = b_int; fa.p_bnd = p_bnd; fa.b_bnd = b_bnd; fa.p_obs = p_obs; fa.b_obs = b_obs;
  fa.dims = D; fa.data_loss = data_loss ? 1 : 0;
  fa.inv_int = a.inv_int; fa.inv_bnd = a.inv_bnd; fa.inv_obs = a.inv_obs;
  fa.inv_all = 1.0f / (float)((double)n_geom * (ni + nb));
  for (int i = 0; i < 16; ++i) fa.weights[i] = prm->weights[i];
  fa.wdev = weights_dev;
  fa.out = out;
  residual_finish_kernel<<<1, 256, 0, st>>>(fa);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

extern "C" int pcfd_residual_loss(const float* data, int32_t n_geom, int64_t n_rows, int32_t f,
                                  const int64_t* internal_ids, int64_t ni, const int64_t* boundary_ids, int64_t nb,
                                  const int64_t* obs_ids, int64_t no, const float* y_int, int64_t y_plane_stride,
                                  const float* y_bnd, int32_t ldy, const pcfd_residual_params_t* prm,
                                  float* gy_int, float* gy_bnd, float* out, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  return pcfd_residual_loss_w(data, n_geom, n_rows, f, internal_ids, ni, boundary_ids, nb, obs_ids, no, y_int,
                              y_plane_stride, y_bnd, ldy, prm, nullptr, nullptr, nullptr, gy_int, gy_bnd, out, workspace,
                              workspace_bytes, stream);
}

// The residual stage of a training step in one kernel (+ one memset node): same results as pcfd_residual_loss_w.
// `ticket`: one int32 that is zero before the first call; the kernel leaves it at zero (the caller keeps it allocated
// and never shares it between streams).  gy_int plane 0 and gy_bnd are zeroed here (cudaMemsetAsync on `stream`).
extern "C" int pcfd_residual_step(const float* data, int32_t n_geom, int64_t n_rows, int32_t f,
                                  const int64_t* internal_ids, int64_t ni, const int64_t* boundary_ids, int64_t nb,
                                  const int64_t* obs_ids, int64_t no, const float* y_int, int64_t y_plane_stride,
                                  const float* y_bnd, int32_t ldy, const pcfd_residual_params_t* prm,
                                  const float* weights_dev, const float* visc_extra, float* gvisc, float* gy_int,
                                  float* gy_bnd, float* out, int32_t* ticket, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  if (!data || !internal_ids || !boundary_ids || !y_int || !y_bnd || !prm || !gy_int || !gy_bnd || !out || !workspace || !ticket)
    return PCFD_ERR_ARG;
  if (n_geom <= 0 || ni <= 0 || nb <= 0 || (prm->dims != 2 && prm->dims != 3) || ldy < prm->dims + 1) return PCFD_ERR_ARG;
  if (workspace_bytes < pcfd_residual_workspace_bytes(n_geom, ni, nb, no)) return PCFD_ERR_WORKSPACE;
  const bool data_loss = prm->enable_data_loss && no > 0;
  if (data_loss && !obs_ids) return PCFD_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  ResArgs a;
  a.data = data; a.n_rows = n_rows; a.f = f;
  a.internal_ids = internal_ids; a.ni = ni; a.boundary_ids = boundary_ids; a.nb = nb; a.obs_ids = obs_ids; a.no = no;
  a.y_int = y_int; a.ps = y_plane_stride; a.y_bnd = y_bnd; a.ldy = ldy; a.gy_int = gy_int; a.gy_bnd = gy_bnd;
  a.p = *prm; a.n_geom = n_geom; a.wdev = weights_dev; a.fields = nullptr; a.visc_extra = visc_extra; a.gvisc = gvisc;
  a.vec_rows = rows_vectorisable(y_int, gy_int, y_plane_stride, ldy);
  a.inv_int = 1.0f / (float)((double)n_geom * ni);
  a.inv_bnd = 1.0f / (float)((double)n_geom * nb);
  a.inv_obs = no > 0 ? 1.0f / (float)((double)n_geom * no) : 0.0f;
  a.partial = nullptr;
  float* ws = reinterpret_cast<float*>(workspace);
  const int b_int = blocks_for((int64_t)n_geom * ni), b_bnd = blocks_for((int64_t)n_geom * nb);
  const int b_obs = data_loss ? blocks_for((int64_t)n_geom * no) : 0;
  FinishArgs fa;
  fa.p_int = ws; fa.b_int = b_int; fa.p_bnd = ws + (size_t)b_int * NSLOT; fa.b_bnd = b_bnd;
  fa.p_obs = fa.p_bnd + (size_t)b_bnd * NSLOT; fa.b_obs = b_obs;
  const int D = prm->dims;
  fa.dims = D; fa.data_loss = data_loss ? 1 : 0;
  fa.inv_int = a.inv_int; fa.inv_bnd = a.inv_bnd; fa.inv_obs = a.inv_obs;
  fa.inv_all = 1.0f / (float)((double)n_geom * (ni + nb));
  for (int i = 0; i < 16; ++i) fa.weights[i] = prm->weights[i];
  fa.wdev = weights_dev;
  fa.out = out;
  // value plane of gy: accumulated with atomics by all three roles.  When the two buffers are adjacent (ops.py allocates
  // them as one block: [gy_bnd | gy_int]) one memset covers both.
  const size_t bytes_int = (size_t)n_geom * ni * ldy * sizeof(float), bytes_bnd = (size_t)n_geom * nb * ldy * sizeof(float);
  cudaError_t e;
  if (reinterpret_cast<char*>(gy_bnd) + bytes_bnd == reinterpret_cast<char*>(gy_int)) {
    e = cudaMemsetAsync(gy_bnd, 0, bytes_bnd + bytes_int, st);
  } else {
    e = cudaMemsetAsync(gy_bnd, 0, bytes_bnd, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(gy_int, 0, bytes_int, st);
  }
  if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;
  const int grid = b_int + b_bnd + b_obs;
  if (D == 2 && prm->lap_mode == PCFD_LAP_REFERENCE) residual_step_kernel<2, PCFD_LAP_REFERENCE><<<grid, 256, 0, st>>>(a, fa, b_int, b_bnd, ticket);
  else if (D == 2) residual_step_kernel<2, PCFD_LAP_TRUE><<<grid, 256, 0, st>>>(a, fa, b_int, b_bnd, ticket);
  else if (prm->lap_mode == PCFD_LAP_REFERENCE) residual_step_kernel<3, PCFD_LAP_REFERENCE><<<grid, 256, 0, st>>>(a, fa, b_int, b_bnd, ticket);
  else residual_step_kernel<3, PCFD_LAP_TRUE><<<grid, 256, 0, st>>>(a, fa, b_int, b_bnd, ticket);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

// Per-point residual map of the internal points (predict_step with verbose_predict, models/model_base.py:233-252):
// fields [n_geom*ni][D+1] = (momentum residual x D, divergence).
extern "C" int pcfd_residual_fields(const float* data, int32_t n_geom, int64_t n_rows, int32_t f,
                                    const int64_t* internal_ids, int64_t ni, const float* y_int, int64_t y_plane_stride,
                                    int32_t ldy, const pcfd_residual_params_t* prm, float* fields, void* stream) {
  if (!data || !internal_ids || !y_int || !prm || !fields) return PCFD_ERR_ARG;
  if (n_geom <= 0 || ni <= 0 || (prm->dims != 2 && prm->dims != 3) || ldy < prm->dims + 1) return PCFD_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  ResArgs a;
  a.data = data; a.n_rows = n_rows; a.f = f;
  a.internal_ids = internal_ids; a.ni = ni; a.boundary_ids = nullptr; a.nb = 0; a.obs_ids = nullptr; a.no = 0;
  a.y_int = y_int; a.ps = y_plane_stride; a.y_bnd = nullptr; a.ldy = ldy; a.gy_int = nullptr; a.gy_bnd = nullptr;
  a.p = *prm; a.n_geom = n_geom; a.wdev = nullptr; a.fields = fields; a.partial = nullptr;
  a.visc_extra = nullptr; a.gvisc = nullptr;
  a.vec_rows = rows_vectorisable(y_int, nullptr, y_plane_stride, ldy);
  a.inv_int = 1.0f / (float)((double)n_geom * ni); a.inv_bnd = 0.0f; a.inv_obs = 0.0f;
  const int b_int = blocks_for((int64_t)n_geom * ni);
  const int D = prm->dims;
  if (D == 2 && prm->lap_mode == PCFD_LAP_REFERENCE) residual_internal_kernel<2, PCFD_LAP_REFERENCE><<<b_int, 256, 0, st>>>(a);
  else if (D == 2) residual_internal_kernel<2, PCFD_LAP_TRUE><<<b_int, 256, 0, st>>>(a);
  else if (prm->lap_mode == PCFD_LAP_REFERENCE) residual_internal_kernel<3, PCFD_LAP_REFERENCE><<<b_int, 256, 0, st>>>(a);
  else residual_internal_kernel<3, PCFD_LAP_TRUE><<<b_int, 256, 0, st>>>(a);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

// ReLoBRaLo loss balancing on the device (models/losses.py:64-124), one thread: reads the unscaled loss terms of the
// current step and the module's buffers, writes the weights the residual pass applies.  `step` counts calls (the
// reference's global_step); rho ~ Bernoulli(beta) comes from a counter-based hash of (seed, step) instead of
// torch.bernoulli's global generator.
__global__ void relobralo_kernel(const float* __restrict__ losses, int n, float* init_l, float* prev_l, float* lam,
                                 int64_t* step, int batch_size, float alpha, float beta, float tau, float eps,
                                 uint64_t seed, float* weights) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int64_t s = *step;
  if (s == 0) {
    for (int i = 0; i < n; ++i) { init_l[i] = losses[i]; prev_l[i] = losses[i]; weights[i] = 1.0f; }
  } else {
    if (s % batch_size == 0) {
      float np = -INFINITY, ni = -INFINITY;
      for (int i = 0; i < n; ++i) {
        prev_l[i] = prev_l[i] / (float)batch_size;
        np = fmaxf(np, losses[i] / (tau * prev_l[i]));
        ni = fmaxf(ni, losses[i] / (tau * init_l[i]));
      }
      const float u = (float)(mix64(seed ^ (uint64_t)s * 0x9e3779b97f4a7c15ULL) >> 40) * (1.0f / 16777216.0f);
      const float rho = u < beta ? 1.0f : 0.0f;
      float lp[16], li[16], sp = 0.0f, si = 0.0f;
      for (int i = 0; i < n; ++i) {
        lp[i] = expf(losses[i] / (tau * prev_l[i] + eps) - np);
        li[i] = expf(losses[i] / (tau * init_l[i] + eps) - ni);
        sp += lp[i]; si += li[i];
      }
      for (int i = 0; i < n; ++i) {
        lp[i] *= (float)n / (sp + eps);
        li[i] *= (float)n / (si + eps);
        lam[i] = alpha * (rho * lam[i] + (1.0f - rho) * li[i]) + (1.0f - alpha) * lp[i];
        prev_l[i] = losses[i];
      }
    } else {
      for (int i = 0; i < n; ++i) prev_l[i] += losses[i];
    }
    for (int i = 0; i < n; ++i) weights[i] = lam[i];
  }
  *step = s + 1;
}

extern "C" int pcfd_relobralo_update(const float* losses, int32_t n, float* init_losses, float* prev_losses,
                                     float* lambda_ema, int64_t* step, int32_t batch_size, float alpha, float beta,
                                     float tau, float eps, uint64_t seed, float* weights_out, void* stream) {
  if (!losses || !init_losses || !prev_losses || !lambda_ema || !step || !weights_out || n <= 0 || n > 16 || batch_size <= 0)
    return PCFD_ERR_ARG;
  relobralo_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(losses, n, init_losses, prev_losses, lambda_ema, step, batch_size, alpha,
                                                       beta, tau, eps, seed, weights_out);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}
