// Engine 2, forward of a VALUE-ONLY jet layer (cj = 1: boundary chain, set-abstraction / encoder MLPs) with the A operand
// in tensor memory.
//
//   zout[row][:] = T(zin)[row][:] * W^T  (+ bias / per-geometry constant)
//
// The general forward kernel (ws_fwd.cu) keeps the transformed activations and their TF32 remainders in shared memory,
// where the tensor core reads them back: with 4-byte operands that traffic saturates the SM's shared-memory pipe
// (DESIGN.md section 3).  With one channel there is no cross-lane coupling in the input transform -- a thread owns one
// point -- so the transform warps can hand the A operand to the tensor core through TMEM instead:
//
//   TMA            raw pre-activation tile [128 points x 16 entries] + weight tile [NT x 16] -> shared memory (8-stage ring)
//   transform      thread = point (TMEM lane): 4 x LDS.128 of its row, activation / dropout / branch scaling in registers,
//                  tcgen05.st of the row (hi = the fp32 word itself, the tensor core reads its top 19 bits) and of the
//                  exact remainder lo = x - trunc_tf32(x) into the stage's TMEM slot; remainder tile of W in shared memory
//   MMA            D += Ahi[tmem]*Bhi + Alo[tmem]*Bhi + Ahi[tmem]*Blo  (tcgen05.mma, A from TMEM, B from shared memory)
//   epilogue       as in ws_fwd.cu (tcgen05.ld -> bias -> swizzled staging -> TMA store), accumulators double-buffered
//
// Shared-memory traffic per 128 x 16 stage: TMA 16 KB + transform 24 KB + B operand reads 24 KB, against 92 KB for the
// same rows in the general kernel; TMEM: 2 x NT accumulator columns + 8 stages x 32 operand columns = 512.
// (A-in-TMEM operand layout validated by scripts/probe/ts_probe.cu: lane = row, one 32-bit column per K entry.)
#include <cstdlib>

#include "common.cuh"
#include "ws_common.cuh"

namespace pcfd {
namespace ws {

constexpr int F1_STAGES = 8;
constexpr int F1_GROUPS = 4;                         // transform groups of 4 warps (one warp per TMEM lane quarter)
constexpr int F1_W_TMA = 4 * F1_GROUPS, F1_W_MMA = F1_W_TMA + 1, F1_W_EPI = F1_W_TMA + 2;   // epilogue warps W_EPI .. W_EPI+3
constexpr int F1_THREADS = (F1_W_EPI + 4) * 32;
constexpr int F1_A_BYTES = 128 * 64;                 // raw A tile of a stage
constexpr uint32_t F1_A_TMEM = 256;                  // first operand column (accumulators use [0, 2*NT))
static_assert(F1_STAGES % F1_GROUPS == 0, "a group must see every use of its stages");

// transform of one point's 16 staged entries (columns col0 .. col0+15, all inside the transformed range): straight-line,
// so that the 16 independent activation / dropout chains interleave
template <int ACT, bool DROP>
__device__ __forceinline__ void f1_row(const uint8_t* st, int point, const InTrans& tin, uint32_t hseed, int64_t row,
                                       int64_t geom, int col0, uint32_t (&hi)[16], uint32_t (&lo)[16]) {
  float x[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 v = *reinterpret_cast<const float4*>(st + swz<64>(point, j));
    x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
  }
  const uint32_t hrow = DROP ? dropout_row_hash(hseed, row) : 0u;
  const float* es = tin.escale != nullptr ? tin.escale + geom * tin.ldescale + col0 : nullptr;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    float sc = 1.0f;
    if (DROP) sc = dropout_from_row(hrow, row, col0 + e, tin.drop_p, tin.inv_keep);
    if (es != nullptr) sc *= __ldg(es + e);
    float zz[1] = {x[e]};
    jet_act_fwd_t<1, ACT>(sc, zz);
    hi[e] = __float_as_uint(zz[0]);
    lo[e] = __float_as_uint(zz[0] - trunc_tf32(zz[0]));
  }
}

struct Fwd1Args {
  const float* bias; const float* cvec; int ldcvec;
  int64_t rows, rows_per_geom; int k, n;
  InTrans tin;
  int row_tiles, n_passes;
  int vec_const;
};

template <int NT>
__global__ void __launch_bounds__(F1_THREADS, 1) ws_fwd1_kernel(const __grid_constant__ CUtensorMap tmZ,
                                                                const __grid_constant__ CUtensorMap tmW,
                                                                const __grid_constant__ CUtensorMap tmO, Fwd1Args a) {
  constexpr int B_BYTES = NT * 64;
  constexpr int STAGE_BYTES = F1_A_BYTES + 2 * B_BYTES;
  constexpr uint32_t TX_BYTES = F1_A_BYTES + B_BYTES;
  static_assert(NT == 64 || NT == 128, "NT must be 64 or 128");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* epi_stage = smem + F1_STAGES * STAGE_BYTES;   // 4 warps x 4 KB
  __shared__ __align__(8) uint64_t raw_full[F1_STAGES], ops_ready[F1_STAGES], stage_free[F1_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_free[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = uniform_warp_id(), lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < F1_STAGES; ++s) {
      tc::mbar_init(&raw_full[s], 1);
      tc::mbar_init(&ops_ready[s], 128);
      tc::mbar_init(&stage_free[s], 1);
    }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&acc_full[b], 1); tc::mbar_init(&acc_free[b], 128); }
    tc::fence_mbar_init();
  }
  if (warp == F1_W_MMA) tc::tmem_alloc(&tmem_base_s, 512);
  if (warp == F1_W_TMA && lane == 0) { prefetch_tmap(&tmZ); prefetch_tmap(&tmW); }
  if (warp == F1_W_EPI && lane == 0) prefetch_tmap(&tmO);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  griddep_wait();                  // everything above overlaps the tail of the previous kernel of the stream
  griddep_launch_dependents();     // one resident wave: the next kernel may take SMs as they free up

  const int total_tiles = a.row_tiles * a.n_passes;
  const int nkc = (a.k + BK - 1) / BK;
  const int my_tiles = total_tiles > (int)blockIdx.x ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == F1_W_TMA) {
    // ================================ TMA producer ================================
    uint32_t it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int rt = t / a.n_passes, np = t - rt * a.n_passes;
      const int row0 = rt * 128;
      for (int kc = 0; kc < nkc; ++kc, ++it) {
        const int s = it % F1_STAGES;
        const uint32_t ph = (it / F1_STAGES) & 1;
        tc::bounded_wait(&stage_free[s], ph ^ 1);
        uint8_t* st = smem + s * STAGE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(&raw_full[s], TX_BYTES);
          tma_load_2d(st, &tmZ, kc * BK, row0, &raw_full[s]);
          tma_load_2d(st + F1_A_BYTES, &tmW, kc * BK, np * NT, &raw_full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == F1_W_MMA) {
    // ================================ MMA issuer ================================
    constexpr uint32_t IDESC = tc::make_idesc_tf32(128, NT, false, false);
    const uint64_t dbase = desc_kmajor<64>(tc::smem_u32(smem));
    uint32_t it = 0, tl = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
      const uint32_t buf = tl & 1, aph = (tl >> 1) & 1;
      tc::bounded_wait(&acc_free[buf], aph ^ 1);
      tc::tc_fence_after();
      const uint32_t d = tmem_base + buf * NT;
      for (int kc = 0; kc < nkc; ++kc, ++it) {
        const int s = it % F1_STAGES;
        const uint32_t ph = (it / F1_STAGES) & 1;
        tc::bounded_wait(&ops_ready[s], ph);
        tc::tc_fence_after();
        const uint64_t ds = dbase + (uint64_t)((uint32_t)(s * STAGE_BYTES + F1_A_BYTES) >> 4);
        const uint32_t ta = tmem_base + F1_A_TMEM + (uint32_t)s * 32;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint64_t db_hi = ds + ((ks * 32) >> 4);
            const uint64_t db_lo = ds + ((B_BYTES + ks * 32) >> 4);
            mma_tf32_ts(d, ta + ks * 8, db_hi, IDESC, (kc > 0 || ks > 0) ? 1u : 0u);
            mma_tf32_ts(d, ta + 16 + ks * 8, db_hi, IDESC, 1u);
            mma_tf32_ts(d, ta + ks * 8, db_lo, IDESC, 1u);
          }
          tc::mma_commit(&stage_free[s]);
        }
        __syncwarp();
      }
      if (elect_one()) tc::mma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else if (warp < F1_W_TMA) {
    // ================================ transform: thread = point = TMEM lane ================================
    const int g = warp >> 2, q = warp & 3;
    const int point = 32 * q + lane;
    const int tt = q * 32 + lane;                              // 0..127 within the group
    const uint64_t seed = a.tin.seed_dev ? *a.tin.seed_dev : 0ULL;
    const uint32_t hseed = dropout_seed_hash(seed, a.tin.salt);
    const bool scaled = a.tin.escale != nullptr || a.tin.drop_p > 0.0f;
    const bool plain = a.tin.act == PCFD_ACT_NONE && !scaled;
    const bool drop = a.tin.drop_p > 0.0f;
    const uint32_t n_it = (uint32_t)my_tiles * (uint32_t)nkc;
    for (uint32_t it = g; it < n_it; it += F1_GROUPS) {
      const int s = (int)(it % F1_STAGES);
      const uint32_t ph = (it / F1_STAGES) & 1;
      const uint32_t tl = it / (uint32_t)nkc;
      const int kc = (int)(it - tl * (uint32_t)nkc);
      const int t = (int)blockIdx.x + (int)tl * (int)gridDim.x;
      int64_t row = (int64_t)(t / a.n_passes) * 128 + point;
      if (row >= a.rows) row = a.rows - 1;   // rows past the end are zero-filled by TMA and clipped by the store
      const int64_t geom = a.tin.escale != nullptr ? geom_of(row, a.rows_per_geom) : 0;
      uint8_t* st = smem + s * STAGE_BYTES;
      uint8_t* bt = st + F1_A_BYTES;
      tc::bounded_wait(&raw_full[s], ph);
      tc::tc_fence_after();   // the MMAs that read this stage's TMEM slot completed before its refill was issued
      // ---- B: remainder tile of the weights
#pragma unroll
      for (int i0 = 0; i0 < NT * 4; i0 += 128) {
        const int item = i0 + tt;
        const float4 x = *reinterpret_cast<const float4*>(bt + swz<64>(item >> 2, item & 3));
        *reinterpret_cast<float4*>(bt + B_BYTES + swz<64>(item >> 2, item & 3)) =
            make_float4(x.x - trunc_tf32(x.x), x.y - trunc_tf32(x.y), x.z - trunc_tf32(x.z), x.w - trunc_tf32(x.w));
      }
      // ---- A: this point's 16 entries -> registers -> transform -> TMEM (hi, lo)
      uint32_t hi[16], lo[16];
      const int col0 = kc * BK;
      if (plain || col0 + BK <= a.tin.act_cols) {
        // whole chunk inside the transformed columns: one straight-line body for the 16 independent entries
        if (plain) f1_row<PCFD_ACT_NONE, false>(st, point, a.tin, hseed, row, geom, col0, hi, lo);
        else if (a.tin.act == PCFD_ACT_SILU) {
          if (drop) f1_row<PCFD_ACT_SILU, true>(st, point, a.tin, hseed, row, geom, col0, hi, lo);
          else f1_row<PCFD_ACT_SILU, false>(st, point, a.tin, hseed, row, geom, col0, hi, lo);
        } else if (a.tin.act == PCFD_ACT_TANH) {
          if (drop) f1_row<PCFD_ACT_TANH, true>(st, point, a.tin, hseed, row, geom, col0, hi, lo);
          else f1_row<PCFD_ACT_TANH, false>(st, point, a.tin, hseed, row, geom, col0, hi, lo);
        } else {
          if (drop) f1_row<PCFD_ACT_NONE, true>(st, point, a.tin, hseed, row, geom, col0, hi, lo);
          else f1_row<PCFD_ACT_NONE, false>(st, point, a.tin, hseed, row, geom, col0, hi, lo);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 x = *reinterpret_cast<const float4*>(st + swz<64>(point, j));
          float v[1][4] = {{x.x, x.y, x.z, x.w}};
          const int c0 = col0 + j * 4;
          if (c0 < a.tin.act_cols) transform_dispatch<1>(v, a.tin, scaled, hseed, row, geom, c0, a.tin.act_cols - c0);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            hi[4 * j + e] = __float_as_uint(v[0][e]);
            lo[4 * j + e] = __float_as_uint(v[0][e] - trunc_tf32(v[0][e]));
          }
        }
      }
      const uint32_t ta = tmem_base + ((uint32_t)(32 * q) << 16) + F1_A_TMEM + (uint32_t)s * 32;
      tmem_st16(ta, hi);
      tmem_st16(ta + 16, lo);
      tmem_st_wait();
      tc::fence_proxy_async();
      tc::tc_fence_before();
      mbar_arrive(&ops_ready[s]);
    }
  } else {
    // ================================ epilogue (warps W_EPI .. W_EPI+3) ================================
    const int q = warp & 3;                           // TMEM lane quarter this warp may read
    uint8_t* sbuf = epi_stage + (warp - F1_W_EPI) * 4096;
    uint8_t* srow = sbuf + lane * 128;
    const uint32_t x7 = (uint32_t)(lane & 7) << 4;
    const bool add_const = a.bias != nullptr || a.cvec != nullptr;
    uint32_t tl = 0;
    bool pending = false;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
      const int rt = t / a.n_passes, np = t - rt * a.n_passes;
      const int64_t row0 = (int64_t)rt * 128;
      const uint32_t buf = tl & 1, aph = (tl >> 1) & 1;
      tc::bounded_wait(&acc_full[buf], aph);
      tc::tc_fence_after();
      const int64_t row = row0 + 32 * q + lane;
      const float* cv = nullptr;
      if (a.cvec != nullptr && row < a.rows) cv = a.cvec + geom_of(row, a.rows_per_geom) * a.ldcvec;
      const uint32_t tcol = tmem_base + ((uint32_t)(32 * q) << 16) + buf * NT;
#pragma unroll 1
      for (int cb = 0; cb < NT / 32; ++cb) {
        const int col0 = np * NT + cb * 32;
        if (col0 >= a.n) break;
        uint32_t r[32];
        tmem_ld32_nowait(tcol + cb * 32, r);
        tmem_ld_wait();
        if (add_const) {
          if (a.vec_const && col0 + 32 <= a.n) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (a.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(a.bias + col0 + i));
              if (cv != nullptr) {
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(cv + col0 + i));
                b4.x += c4.x; b4.y += c4.y; b4.z += c4.z; b4.w += c4.w;
              }
              r[i] = __float_as_uint(__uint_as_float(r[i]) + b4.x);
              r[i + 1] = __float_as_uint(__uint_as_float(r[i + 1]) + b4.y);
              r[i + 2] = __float_as_uint(__uint_as_float(r[i + 2]) + b4.z);
              r[i + 3] = __float_as_uint(__uint_as_float(r[i + 3]) + b4.w);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (col0 + i < a.n) {
                float b = a.bias != nullptr ? __ldg(a.bias + col0 + i) : 0.0f;
                if (cv != nullptr) b += __ldg(cv + col0 + i);
                r[i] = __float_as_uint(__uint_as_float(r[i]) + b);
              }
            }
          }
        }
        if (pending) {
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(srow + (((uint32_t)j << 4) ^ x7)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmO, sbuf, col0, (int)(row0 + 32 * q));
          tma_store_commit();
        }
        pending = true;
      }
      tc::tc_fence_before();
      mbar_arrive(&acc_free[buf]);
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == F1_W_MMA) tc::tmem_dealloc(tmem_base, 512);
}

template <int NT>
static int launch_fwd1(const float* zin, int ldzin, const float* w, int ldw, float* zout, int ldzout, Fwd1Args a,
                       cudaStream_t st) {
  constexpr int SMEM = F1_STAGES * (F1_A_BYTES + 2 * NT * 64) + 4 * 4096 + 1024;
  CUtensorMap tmZ, tmW, tmO;
  {
    const uint64_t dims[2] = {(uint64_t)a.k, (uint64_t)a.rows};
    const uint64_t str[1] = {(uint64_t)ldzin * 4};
    const uint32_t box[2] = {BK, 128};
    if (!make_tmap(&tmZ, zin, 2, dims, str, box, 64)) return PCFD_ERR_ARG;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.k, (uint64_t)a.n};
    const uint64_t str[1] = {(uint64_t)ldw * 4};
    const uint32_t box[2] = {BK, NT};
    if (!make_tmap(&tmW, w, 2, dims, str, box, 64)) return PCFD_ERR_ARG;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.n, (uint64_t)a.rows};
    const uint64_t str[1] = {(uint64_t)ldzout * 4};
    const uint32_t box[2] = {32, 32};
    if (!make_tmap(&tmO, zout, 2, dims, str, box, 128)) return PCFD_ERR_ARG;
  }
  a.vec_const = (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.cvec) & 15) == 0 && a.ldcvec % 4 == 0;
  a.row_tiles = (int)((a.rows + 127) / 128);
  a.n_passes = (a.n + NT - 1) / NT;
  {
    const cudaError_t e = ensure_dyn_smem<ws_fwd1_kernel<NT>>(SMEM);
    if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;
  }
  const int total = a.row_tiles * a.n_passes;
  const int grid = balanced_grid(total);
  const cudaError_t le = launch_pdl(ws_fwd1_kernel<NT>, dim3(grid), dim3(F1_THREADS), (size_t)SMEM, st, tmZ, tmW, tmO, a);
  if (le != cudaSuccess) return PCFD_ERR_CUDA + (int)le;
  return PCFD_OK;
}

}  // namespace ws
}  // namespace pcfd

using namespace pcfd;

// value-only forward with the A operand in tensor memory (same preconditions as pcfd_ws_supported_fwd, cj = 1);
// PCFD_FWD1=0 keeps such layers on the general kernel
// ... and layers with fewer than PCFD_FWD1_MIN_K (default 32) input columns, whose single ring stage leaves the
// time in the epilogue, where the general kernel has twice the warps
extern "C" int pcfd_ws_fwd1_enabled(int32_t k) {
  static int on = -1, min_k = 32;
  if (on < 0) {
    const char* m = getenv("PCFD_FWD1_MIN_K");
    if (m) min_k = atoi(m);
    const char* e = getenv("PCFD_FWD1");
    on = e ? atoi(e) : 1;
  }
  return on && k >= min_k;
}

extern "C" int pcfd_ws_jet_linear_fwd1(const float* zin, int32_t ldzin, const pcfd_intrans_t* tin, const float* w,
                                       int32_t ldw, const float* bias, const float* cvec, int32_t ldcvec, float* zout,
                                       int32_t ldzout, int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n,
                                       void* stream) {
  ws::Fwd1Args a{bias, cvec, ldcvec, rows, rows_per_geom, k, n, make_intrans(tin, k), 0, 0, 0};
  cudaStream_t st = (cudaStream_t)stream;
  return n <= 64 ? ws::launch_fwd1<64>(zin, ldzin, w, ldw, zout, ldzout, a, st)
                 : ws::launch_fwd1<128>(zin, ldzin, w, ldw, zout, ldzout, a, st);
}
