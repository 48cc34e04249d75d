// Engine 2, forward jet layer: persistent warp-specialised tcgen05 kernel.
//
//   zout[c][row][:] = T(zin)[c][row][:] * W^T  (+ bias / per-geometry constant on channel 0)
//
// One CTA per SM walks (row tile, output-column pass) pairs.  A row tile is 8 "slabs" of 32 points of
// one channel each (2 UMMA sub-tiles of 128 lanes): CJ = 4 -> 2 point groups x 4 channels, CJ = 1 ->
// 8 point groups, CJ = 3 -> 2 x 3 (+2 idle slabs), CJ = 5 / 7 -> 1 x CJ.  All channels of a point sit
// in the same stage, so the activation jet can be applied to the staged tile in place.
//
//   warp 16     TMA producer: raw fp32 tiles of the pre-activation jets (3-D map, 64B swizzle) and of
//               the weights into a 4-stage ring (16 contraction entries per stage)
//   warps 0-15  transform, NGROUPS groups (2 groups of 8 warps: measured best of 4x4 / 2x8 / 1x16), group g takes ring
//               iterations g, g + NGROUPS, ...: activation jet / dropout / branch scaling in shared memory,
//               in place (the tensor core reads the top 19 bits of an fp32 word, so the transformed fp32
//               tile IS the TF32 "hi" operand) + the exact remainder lo = x - trunc_tf32(x) into a
//               second tile
//   warp 17     MMA issuer: D += Ahi*Bhi + Alo*Bhi + Ahi*Blo (3xTF32), accumulators double-buffered in
//               TMEM (2 buffers x 2 sub-tiles x NT columns)
//   warps 18-25 epilogue (one warp per TMEM lane quarter and sub-tile; the highest warp ids, which the
//               issue arbiter favours): tcgen05.ld -> bias -> swizzled staging tile -> TMA store,
//               overlapping the next tile's main loop
#include <cstdlib>

#include "common.cuh"
#include "ws_common.cuh"

namespace pcfd {
namespace ws {

#ifndef PCFD_FWD_GROUPS
#define PCFD_FWD_GROUPS 2
#endif
constexpr int NGROUPS = PCFD_FWD_GROUPS;   // transform groups; group g takes ring iterations g, g + NGROUPS, ... (NGROUPS divides STAGES)
constexpr int GROUP = 512 / NGROUPS;       // threads of one transform group
constexpr int W_TMA = 16, W_MMA = W_TMA + 1, W_EPI = W_TMA + 2;
static_assert(STAGES % NGROUPS == 0, "a group must see every use of its stages");
constexpr int THREADS = (W_EPI + 8) * 32;

struct FwdArgs {
  float* zout; int64_t zout_ps; int ldzout;
  const float* bias; const float* cvec; int ldcvec;
  int64_t rows, rows_per_geom; int k, n;
  InTrans tin;
  int row_tiles, n_passes;
  int vec_const;   // bias / cvec rows are 16-byte aligned
  int dbg;   // PCFD_WS_DEBUG bit mask (timing experiments): 1 skip transform, 2 skip epilogue, 4 skip MMAs
};

template <int CJ, int NT>
__global__ void __launch_bounds__(THREADS, 1) ws_fwd_kernel(const __grid_constant__ CUtensorMap tmZ,
                                                            const __grid_constant__ CUtensorMap tmW,
                                                            const __grid_constant__ CUtensorMap tmO, FwdArgs a) {
  constexpr int PGS = 8 / CJ;                       // point groups per tile
  constexpr int SLABS = PGS * CJ;                   // slabs in use (of 8)
  constexpr int POINTS = 32 * PGS;
  constexpr int B_BYTES = NT * 64;
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr uint32_t TMEM_COLS = 4 * NT;            // 2 buffers x 2 sub-tiles x NT
  constexpr uint32_t TX_BYTES = SLABS * SLAB_BYTES + B_BYTES;
  static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "NT must be 64 or 128");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* epi_stage = smem + STAGES * STAGE_BYTES;   // 8 warps x 4 KB
  __shared__ __align__(8) uint64_t raw_full[STAGES], ops_ready[STAGES], stage_free[STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_free[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = uniform_warp_id(), lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(&raw_full[s], 1);
      tc::mbar_init(&ops_ready[s], GROUP);
      tc::mbar_init(&stage_free[s], 1);
    }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&acc_full[b], 1); tc::mbar_init(&acc_free[b], 256); }
    tc::fence_mbar_init();
  }
  if (warp == W_MMA) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);
  if (warp == W_TMA && lane == 0) { prefetch_tmap(&tmZ); prefetch_tmap(&tmW); }
  if (warp == W_EPI && lane == 0) prefetch_tmap(&tmO);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  griddep_wait();                  // everything above overlaps the tail of the previous kernel of the stream
  griddep_launch_dependents();     // one resident wave: the next kernel may take SMs as they free up

  const int total_tiles = a.row_tiles * a.n_passes;
  const int nkc = (a.k + BK - 1) / BK;
  const int my_tiles = total_tiles > (int)blockIdx.x ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == W_TMA) {
    // ================================ TMA producer ================================
    uint32_t it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int rt = t / a.n_passes, np = t - rt * a.n_passes;
      const int row0 = rt * POINTS;
      for (int kc = 0; kc < nkc; ++kc, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        tc::bounded_wait(&stage_free[s], ph ^ 1);
        uint8_t* st = smem + s * STAGE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(&raw_full[s], TX_BYTES);
          if (CJ == 1) {
            tma_load_3d(st, &tmZ, kc * BK, row0, 0, &raw_full[s]);
          } else {
#pragma unroll
            for (int pg = 0; pg < PGS; ++pg)
              tma_load_3d(st + pg * CJ * SLAB_BYTES, &tmZ, kc * BK, row0 + pg * 32, 0, &raw_full[s]);
          }
          tma_load_2d(st + 2 * A_BYTES, &tmW, kc * BK, np * NT, &raw_full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == W_MMA) {
    // ================================ MMA issuer ================================
    constexpr uint32_t IDESC = tc::make_idesc_tf32(128, NT, false, false);
    const uint64_t dbase = desc_kmajor<64>(tc::smem_u32(smem));      // + (byte offset >> 4) in the address field
    uint32_t it = 0, tl = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
      const uint32_t buf = tl & 1, aph = (tl >> 1) & 1;
      tc::bounded_wait(&acc_free[buf], aph ^ 1);
      tc::tc_fence_after();
      for (int kc = 0; kc < nkc; ++kc, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        tc::bounded_wait(&ops_ready[s], ph);
        tc::tc_fence_after();
        const uint64_t ds = dbase + (uint64_t)((uint32_t)(s * STAGE_BYTES) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint64_t db_hi = ds + ((2 * A_BYTES + ks * 32) >> 4);
            const uint64_t db_lo = ds + ((2 * A_BYTES + B_BYTES + ks * 32) >> 4);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const uint64_t da_hi = ds + ((u * 4 * SLAB_BYTES + ks * 32) >> 4);
              const uint64_t da_lo = ds + ((A_BYTES + u * 4 * SLAB_BYTES + ks * 32) >> 4);
              const uint32_t d = tmem_base + buf * (2 * NT) + u * NT;
              tc::mma_tf32(d, da_hi, db_hi, IDESC, (kc > 0 || ks > 0) ? 1u : 0u);
              tc::mma_tf32(d, da_lo, db_hi, IDESC, 1u);
              tc::mma_tf32(d, da_hi, db_lo, IDESC, 1u);
            }
          }
          tc::mma_commit(&stage_free[s]);
        }
        __syncwarp();
      }
      if (elect_one()) tc::mma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else if (warp < W_TMA) {
    // ================================ transform ================================
    const int g = warp / (GROUP / 32);
    const int tt = tid - g * GROUP;
    const uint64_t seed = a.tin.seed_dev ? *a.tin.seed_dev : 0ULL;
    const uint32_t hseed = dropout_seed_hash(seed, a.tin.salt);
    const bool scaled = a.tin.escale != nullptr || a.tin.drop_p > 0.0f;
    const bool plain = a.tin.act == PCFD_ACT_NONE && !scaled;
    constexpr int A_ITEMS = POINTS * 4;                       // (point, 16-byte chunk) positions, all channels each
    constexpr int A_PER = (A_ITEMS + GROUP - 1) / GROUP;
    constexpr int B_PER = (NT * 4 + GROUP - 1) / GROUP;
    const uint32_t n_it = (uint32_t)my_tiles * (uint32_t)nkc;
    for (uint32_t it = g; it < n_it; it += NGROUPS) {
      const int sidx = (int)(it % STAGES);
      uint8_t* st = smem + sidx * STAGE_BYTES;
      uint8_t* bt = st + 2 * A_BYTES;
      const uint32_t tl = it / (uint32_t)nkc;
      const int kc = (int)(it - tl * (uint32_t)nkc);
      const int t = (int)blockIdx.x + (int)tl * (int)gridDim.x;
      const int64_t row0 = (int64_t)(t / a.n_passes) * POINTS;
      const uint32_t ph = (it / STAGES) & 1;
      tc::bounded_wait(&raw_full[sidx], ph);
      if (!(a.dbg & 1)) {
        // ---- B: remainder tile of the weights
        {
          float4 wv[B_PER];
#pragma unroll
          for (int q = 0; q < B_PER; ++q) {
            const int item = tt + q * GROUP;
            if (item < NT * 4) wv[q] = *reinterpret_cast<const float4*>(bt + swz<64>(item >> 2, item & 3));
          }
#pragma unroll
          for (int q = 0; q < B_PER; ++q) {
            const int item = tt + q * GROUP;
            if (item >= NT * 4) continue;
            const float4 x = wv[q];
            *reinterpret_cast<float4*>(bt + B_BYTES + swz<64>(item >> 2, item & 3)) =
                make_float4(x.x - trunc_tf32(x.x), x.y - trunc_tf32(x.y), x.z - trunc_tf32(x.z), x.w - trunc_tf32(x.w));
          }
        }
        // ---- A: activation jet in place + remainder tile
#pragma unroll
        for (int q = 0; q < A_PER; ++q) {
          const int item = tt + q * GROUP;
          if (item >= A_ITEMS) continue;
          const int point = item >> 2, j = item & 3;
          uint8_t* base = st + (point >> 5) * CJ * SLAB_BYTES + swz<64>(point & 31, j);
          float v[CJ][4];
#pragma unroll
          for (int c = 0; c < CJ; ++c) {
            const float4 x = *reinterpret_cast<const float4*>(base + c * SLAB_BYTES);
            v[c][0] = x.x; v[c][1] = x.y; v[c][2] = x.z; v[c][3] = x.w;
          }
          const int64_t row = row0 + point;
          const int col0 = kc * BK + j * 4;
          if (!plain && row < a.rows && col0 < a.tin.act_cols) {
            const int64_t geom = a.tin.escale != nullptr ? geom_of(row, a.rows_per_geom) : 0;
            transform_dispatch<CJ>(v, a.tin, scaled, hseed, row, geom, col0, a.tin.act_cols - col0);
#pragma unroll
            for (int c = 0; c < CJ; ++c)
              *reinterpret_cast<float4*>(base + c * SLAB_BYTES) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
          }
#pragma unroll
          for (int c = 0; c < CJ; ++c)
            *reinterpret_cast<float4*>(base + A_BYTES + c * SLAB_BYTES) =
                make_float4(v[c][0] - trunc_tf32(v[c][0]), v[c][1] - trunc_tf32(v[c][1]), v[c][2] - trunc_tf32(v[c][2]),
                            v[c][3] - trunc_tf32(v[c][3]));
        }
      }
      tc::fence_proxy_async();
      mbar_arrive(&ops_ready[sidx]);
    }
  } else {
    // ================================ epilogue (warps W_EPI .. W_EPI+7) ================================
    const int q = warp & 3;                           // TMEM lane quarter this warp may read
    const int u = (warp - W_EPI) >> 2;                // sub-tile
    const int slab = 4 * u + q;
    const int pg = slab / CJ, c = slab - pg * CJ;
    uint8_t* sbuf = epi_stage + (warp - W_EPI) * 4096;
    uint8_t* srow = sbuf + lane * 128;
    const uint32_t x7 = (uint32_t)(lane & 7) << 4;
    const bool add_const = c == 0 && (a.bias != nullptr || a.cvec != nullptr);
    uint32_t tl = 0;
    bool pending = false;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
      const int rt = t / a.n_passes, np = t - rt * a.n_passes;
      const int64_t row0 = (int64_t)rt * POINTS;
      const uint32_t buf = tl & 1, aph = (tl >> 1) & 1;
      tc::bounded_wait(&acc_full[buf], aph);
      tc::tc_fence_after();
      if (slab < SLABS && !(a.dbg & 2)) {
        const int64_t row = row0 + pg * 32 + lane;
        const float* cv = nullptr;
        if (a.cvec != nullptr && c == 0 && row < a.rows) cv = a.cvec + geom_of(row, a.rows_per_geom) * a.ldcvec;
        const uint32_t tcol = tmem_base + ((uint32_t)(32 * q) << 16) + buf * (2 * NT) + u * NT;
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
          // staging tile -> TMA store (clipped at the tensor bounds by the hardware)
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
            tma_store_3d(&tmO, sbuf, col0, (int)(row0 + pg * 32), c);
            tma_store_commit();
          }
          pending = true;
        }
      }
      tc::tc_fence_before();
      mbar_arrive(&acc_free[buf]);
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int CJ, int NT>
static int launch_fwd(const float* zin, int64_t zin_ps, int ldzin, const float* w, int ldw, float* zout, int64_t zout_ps,
                      int ldzout, FwdArgs a, cudaStream_t st) {
  constexpr int PGS = 8 / CJ;
  constexpr int POINTS = 32 * PGS;
  constexpr int SMEM = STAGES * (2 * A_BYTES + 2 * NT * 64) + 8 * 4096 + 1024;
  CUtensorMap tmZ, tmW, tmO;
  {
    const uint64_t dims[3] = {(uint64_t)a.k, (uint64_t)a.rows, (uint64_t)CJ};
    const uint64_t str[2] = {(uint64_t)ldzin * 4, (uint64_t)zin_ps * 4};
    const uint32_t box[3] = {BK, CJ == 1 ? 256u : 32u, (uint32_t)CJ};
    if (!make_tmap(&tmZ, zin, 3, dims, str, box, 64)) return PCFD_ERR_ARG;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.k, (uint64_t)a.n};
    const uint64_t str[1] = {(uint64_t)ldw * 4};
    const uint32_t box[2] = {BK, NT};
    if (!make_tmap(&tmW, w, 2, dims, str, box, 64)) return PCFD_ERR_ARG;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a.n, (uint64_t)a.rows, (uint64_t)CJ};
    const uint64_t str[2] = {(uint64_t)ldzout * 4, (uint64_t)zout_ps * 4};
    const uint32_t box[3] = {32, 32, 1};
    if (!make_tmap(&tmO, zout, 3, dims, str, box, 128)) return PCFD_ERR_ARG;
  }
  a.zout = zout; a.zout_ps = zout_ps; a.ldzout = ldzout;
  a.vec_const = (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.cvec) & 15) == 0 && a.ldcvec % 4 == 0;
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("PCFD_WS_DEBUG"); dbg = e ? atoi(e) : 0; }
  a.dbg = dbg;
  a.row_tiles = (int)((a.rows + POINTS - 1) / POINTS);
  a.n_passes = (a.n + NT - 1) / NT;
  {
    const cudaError_t e = ensure_dyn_smem<ws_fwd_kernel<CJ, NT>>(SMEM);
    if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;
  }
  const int total = a.row_tiles * a.n_passes;
  const int grid = balanced_grid(total);
  const cudaError_t le = launch_pdl(ws_fwd_kernel<CJ, NT>, dim3(grid), dim3(THREADS), (size_t)SMEM, st, tmZ, tmW, tmO, a);
  if (le != cudaSuccess) return PCFD_ERR_CUDA + (int)le;
  return PCFD_OK;
}

}  // namespace ws
}  // namespace pcfd

using namespace pcfd;

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int pcfd_ws_supported_fwd(const float* zin, int64_t zin_ps, int32_t ldzin, const float* w, int32_t ldw,
                                     const float* zout, int64_t zout_ps, int32_t ldzout, int32_t cj, int64_t rows,
                                     int32_t k, int32_t n) {
  if (!valid_cj(cj) || rows < 256 || k < 8 || n < 16) return 0;
  if (rows >= (int64_t)1 << 31) return 0;
  if (!al16(zin) || !al16(w) || !al16(zout)) return 0;
  if (ldzin % 4 || ldw % 4 || ldzout % 4) return 0;
  if (cj > 1 && (zin_ps % 4 || zout_ps % 4)) return 0;
  return ws::encode_fn() != nullptr;
}

extern "C" int pcfd_ws_jet_linear_fwd(const float* zin, int64_t zin_ps, int32_t ldzin, const pcfd_intrans_t* tin,
                                      const float* w, int32_t ldw, const float* bias, const float* cvec,
                                      int32_t ldcvec, float* zout, int64_t zout_ps, int32_t ldzout, int32_t cj,
                                      int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n, void* stream) {
  ws::FwdArgs a{nullptr, 0, 0, bias, cvec, ldcvec, rows, rows_per_geom, k, n, make_intrans(tin, k), 0, 0, 0, 0};
  cudaStream_t st = (cudaStream_t)stream;
  if (cj == 1) { zin_ps = (int64_t)rows * ldzin; zout_ps = (int64_t)rows * ldzout; }
  // few row tiles (per-geometry layers): 64-column passes double the number of CTAs; otherwise 128-column tiles
  const int pts = 32 * (8 / cj);
  const int64_t items128 = ((rows + pts - 1) / pts) * ((n + 127) / 128);
  const bool narrow = n <= 64 || items128 * 2 <= ws::num_sms();
#define PCFD_WS_FWD(CJ_)                                                                                          \
  return narrow ? ws::launch_fwd<CJ_, 64>(zin, zin_ps, ldzin, w, ldw, zout, zout_ps, ldzout, a, st)              \
                 : ws::launch_fwd<CJ_, 128>(zin, zin_ps, ldzin, w, ldw, zout, zout_ps, ldzout, a, st);
  switch (cj) {
    case 1: PCFD_WS_FWD(1)
    case 3: PCFD_WS_FWD(3)
    case 4: PCFD_WS_FWD(4)
    case 5: PCFD_WS_FWD(5)
    case 7: PCFD_WS_FWD(7)
  }
#undef PCFD_WS_FWD
  return PCFD_ERR_ARG;
}
