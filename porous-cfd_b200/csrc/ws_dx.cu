// Engine 2, backward of a jet layer to its input: persistent warp-specialised tcgen05 kernel.
//
//   gzin[c][row][:] = reverse_T( gzout[c][row][:] * W ; zin )         (+ d loss / d branch scaling)
//
// Same row tiling as the forward kernel (8 slabs of 32 points x channel, 2 UMMA sub-tiles).  The
// contraction runs over the layer's OUTPUT columns, so the weight tile W[n0..n0+16][kk0..kk0+NT] is an
// MN-major B operand: TMA brings 32-column blocks of W rows in the 32-byte-atom 128B swizzle that
// tcgen05 requires for transposed fp32 operands -- no transposed copy of the weights is ever made.
//
//   warps 0-3    operand remainders (lo = x - trunc_tf32(x)) of the staged gzout / W tiles, warp g owns ring
//                stage g
//   warp 4       TMA producer, warp 5 MMA issuer (3xTF32, TMEM double-buffered)
//   warps 6-21   epilogue.  The reverse activation jet mixes the channels of a point, which live in
//                different TMEM lanes, so each 32-column block goes through a shared staging tile:
//                phase 1 (lane = row) dumps the accumulators, phase 2 (8 lanes = one 128-byte row
//                segment, all channels per thread) applies the reverse jet against coalesced loads of
//                zin and writes gzin with coalesced 16-byte stores.  Layers without an input transform
//                skip phase 2 and TMA-store the staging tile.
#include <cstdlib>

#include "common.cuh"
#include "ws_common.cuh"

namespace pcfd {
namespace ws {

#ifndef PCFD_DX_GROUPS
#define PCFD_DX_GROUPS 4
#endif
constexpr int DX_NGROUPS = PCFD_DX_GROUPS;       // remainder (lo-split) groups; group g takes ring iterations g, g + NGROUPS, ...
constexpr int DX_GROUP = 128 / DX_NGROUPS;      // threads per group (4 warps in total)
constexpr int DX_W_TMA = 4, DX_W_MMA = DX_W_TMA + 1, DX_W_EPI = DX_W_TMA + 2;
static_assert(STAGES % DX_NGROUPS == 0, "a group must see every use of its stages");
constexpr int DX_EPI_WARPS = 16, DX_EPI_THREADS = DX_EPI_WARPS * 32;
constexpr int DX_THREADS = (DX_W_EPI + DX_EPI_WARPS) * 32;

// reverse activation jet of one (row, 4 columns) item, all channels; returns nothing, accumulates d/d escale
template <int CJ, int ACT, bool SCALED>
__device__ __forceinline__ void reverse_chunk(float (&g)[CJ][4], const float4 (&z)[CJ], const InTrans& tin, uint32_t hseed,
                                              int64_t row, int64_t geom, int col0, int ncols, float (&ge)[4]) {
  uint32_t hrow = 0;
  if (SCALED && tin.drop_p > 0.0f) hrow = dropout_row_hash(hseed, row);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (e < ncols) {
      float m = 1.0f, sc = 1.0f;
      if (SCALED) {
        if (tin.drop_p > 0.0f) m = dropout_from_row(hrow, row, col0 + e, tin.drop_p, tin.inv_keep);
        sc = m;
        if (tin.escale != nullptr) sc *= __ldg(tin.escale + geom * tin.ldescale + col0 + e);
      }
      float gg[CJ], zz[CJ];
#pragma unroll
      for (int c = 0; c < CJ; ++c) {
        gg[c] = g[c][e];
        zz[c] = e == 0 ? z[c].x : (e == 1 ? z[c].y : (e == 2 ? z[c].z : z[c].w));
      }
      ge[e] = jet_act_bwd_t<CJ, ACT>(sc, m, zz, gg);
#pragma unroll
      for (int c = 0; c < CJ; ++c) g[c][e] = gg[c];
    }
  }
}
template <int CJ>
__device__ __forceinline__ void reverse_dispatch(float (&g)[CJ][4], const float4 (&z)[CJ], const InTrans& tin, bool scaled,
                                                 uint32_t hseed, int64_t row, int64_t geom, int col0, int ncols,
                                                 float (&ge)[4]) {
  if (tin.act == PCFD_ACT_SILU) {
    if (scaled) reverse_chunk<CJ, PCFD_ACT_SILU, true>(g, z, tin, hseed, row, geom, col0, ncols, ge);
    else reverse_chunk<CJ, PCFD_ACT_SILU, false>(g, z, tin, hseed, row, geom, col0, ncols, ge);
  } else if (tin.act == PCFD_ACT_TANH) {
    if (scaled) reverse_chunk<CJ, PCFD_ACT_TANH, true>(g, z, tin, hseed, row, geom, col0, ncols, ge);
    else reverse_chunk<CJ, PCFD_ACT_TANH, false>(g, z, tin, hseed, row, geom, col0, ncols, ge);
  } else {
    reverse_chunk<CJ, PCFD_ACT_NONE, true>(g, z, tin, hseed, row, geom, col0, ncols, ge);
  }
}

struct DxArgs {
  const float* zin; int64_t zin_ps; int ldzin;
  float* gzin; int64_t gzin_ps; int ldgzin;
  float* gescale; int ldgescale;
  int64_t rows, rows_per_geom; int k, n;
  InTrans tin;
  int row_tiles, k_passes;
  int dbg;   // PCFD_WS_DEBUG bit mask (timing experiments): 8 skip the reverse-jet math, 16 skip phase 2, 32 skip the gzin stores
};

template <int CJ, int NT>
__global__ void __launch_bounds__(DX_THREADS, 1) ws_dx_kernel(const __grid_constant__ CUtensorMap tmG,
                                                              const __grid_constant__ CUtensorMap tmW,
                                                              const __grid_constant__ CUtensorMap tmO, DxArgs a) {
  constexpr int PGS = 8 / CJ;
  constexpr int SLABS = PGS * CJ;
  constexpr int POINTS = 32 * PGS;
  constexpr int B_BYTES = NT * 64;
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr uint32_t TMEM_COLS = 4 * NT;
  constexpr uint32_t TX_BYTES = SLABS * SLAB_BYTES + B_BYTES;
  static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "NT must be 64 or 128");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stg = smem + STAGES * STAGE_BYTES;          // 8 slabs x [32 rows x 128 B]
  __shared__ __align__(8) uint64_t raw_full[STAGES], ops_ready[STAGES], stage_free[STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_free[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float ge_acc[128];

  const int tid = threadIdx.x, warp = uniform_warp_id(), lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(&raw_full[s], 1);
      tc::mbar_init(&ops_ready[s], DX_GROUP);
      tc::mbar_init(&stage_free[s], 1);
    }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&acc_full[b], 1); tc::mbar_init(&acc_free[b], DX_EPI_THREADS); }
    tc::fence_mbar_init();
  }
  if (tid < 128) ge_acc[tid] = 0.0f;
  if (warp == DX_W_MMA) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);
  if (warp == DX_W_TMA && lane == 0) { prefetch_tmap(&tmG); prefetch_tmap(&tmW); }
  if (warp == DX_W_EPI && lane == 0) prefetch_tmap(&tmO);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  griddep_wait();                  // everything above overlaps the tail of the previous kernel of the stream
  griddep_launch_dependents();     // one resident wave: the next kernel may take SMs as they free up

  const int total_tiles = a.row_tiles * a.k_passes;
  const int nkc = (a.n + BK - 1) / BK;                  // contraction chunks (over the layer's outputs)
  const int my_tiles = total_tiles > (int)blockIdx.x ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == DX_W_TMA) {
    // ================================ TMA producer ================================
    uint32_t it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int rt = t / a.k_passes, kp = t - rt * a.k_passes;
      const int row0 = rt * POINTS;
      for (int kc = 0; kc < nkc; ++kc, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        tc::bounded_wait(&stage_free[s], ph ^ 1);
        uint8_t* st = smem + s * STAGE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(&raw_full[s], TX_BYTES);
          if (CJ == 1) {
            tma_load_3d(st, &tmG, kc * BK, row0, 0, &raw_full[s]);
          } else {
#pragma unroll
            for (int pg = 0; pg < PGS; ++pg)
              tma_load_3d(st + pg * CJ * SLAB_BYTES, &tmG, kc * BK, row0 + pg * 32, 0, &raw_full[s]);
          }
#pragma unroll
          for (int cbk = 0; cbk < NT / 32; ++cbk)
            tma_load_2d(st + 2 * A_BYTES + cbk * (BK * 128), &tmW, kp * NT + cbk * 32, kc * BK, &raw_full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == DX_W_MMA) {
    // ================================ MMA issuer ================================
    constexpr uint32_t IDESC = tc::make_idesc_tf32(128, NT, false, true);     // A K-major, B MN-major
    const uint64_t abase = desc_kmajor<64>(tc::smem_u32(smem));              // + (byte offset >> 4) in the address field
    const uint64_t bbase = desc_mnmajor(tc::smem_u32(smem), BK * 128, 512);
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
        const uint64_t so = (uint64_t)((uint32_t)(s * STAGE_BYTES) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint64_t db_hi = bbase + so + ((2 * A_BYTES + ks * 1024) >> 4);
            const uint64_t db_lo = bbase + so + ((2 * A_BYTES + B_BYTES + ks * 1024) >> 4);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const uint64_t da_hi = abase + so + ((u * 4 * SLAB_BYTES + ks * 32) >> 4);
              const uint64_t da_lo = abase + so + ((A_BYTES + u * 4 * SLAB_BYTES + ks * 32) >> 4);
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
  } else if (warp < DX_W_TMA) {
    // ================================ operand remainders ================================
    const int g = warp / (DX_GROUP / 32);
    const int tt = tid - g * DX_GROUP;
    constexpr int A_CH = SLABS * SLAB_BYTES / 16, B_CH = B_BYTES / 16;
    static_assert(A_CH % DX_GROUP == 0 && B_CH % DX_GROUP == 0, "chunks must divide over the group");
    const uint32_t n_it = (uint32_t)my_tiles * (uint32_t)nkc;
    for (uint32_t it = g; it < n_it; it += DX_NGROUPS) {
      const int sidx = (int)(it % STAGES);
      uint8_t* st = smem + sidx * STAGE_BYTES;
      const uint32_t ph = (it / STAGES) & 1;
      tc::bounded_wait(&raw_full[sidx], ph);
#pragma unroll
      for (int q0 = 0; q0 < A_CH / DX_GROUP; q0 += 4) {
        float4 x[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q0 + q < A_CH / DX_GROUP) x[q] = *reinterpret_cast<const float4*>(st + (tt + (q0 + q) * DX_GROUP) * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q0 + q < A_CH / DX_GROUP)
            *reinterpret_cast<float4*>(st + A_BYTES + (tt + (q0 + q) * DX_GROUP) * 16) =
                make_float4(x[q].x - trunc_tf32(x[q].x), x[q].y - trunc_tf32(x[q].y), x[q].z - trunc_tf32(x[q].z),
                            x[q].w - trunc_tf32(x[q].w));
      }
      uint8_t* bt = st + 2 * A_BYTES;
#pragma unroll
      for (int q0 = 0; q0 < B_CH / DX_GROUP; q0 += 4) {
        float4 x[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = *reinterpret_cast<const float4*>(bt + (tt + (q0 + q) * DX_GROUP) * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(bt + B_BYTES + (tt + (q0 + q) * DX_GROUP) * 16) =
              make_float4(x[q].x - trunc_tf32(x[q].x), x[q].y - trunc_tf32(x[q].y), x[q].z - trunc_tf32(x[q].z),
                          x[q].w - trunc_tf32(x[q].w));
      }
      tc::fence_proxy_async();
      mbar_arrive(&ops_ready[sidx]);
    }
  } else {
    // ================================ epilogue (16 warps) ================================
    const int ew = warp - DX_W_EPI;                     // 0..15
    const int q = warp & 3;                             // TMEM lane quarter this warp may read
    const int u = (ew >> 2) & 1;                        // sub-tile
    const int h = ew >> 3;                              // column half of a 32-column block (phase 1)
    const int slab = 4 * u + q;
    const int et = tid - DX_W_EPI * 32;                 // 0..511
    const int j = et & 7, r = (et >> 3) & 31, pgsel = et >> 8;   // phase 2: 16-byte column chunk, row, point-group parity
    constexpr int ITEMS = (PGS + 1) / 2;                // (row, 4 columns) items per thread and block
    const uint32_t x7 = (uint32_t)(lane & 7) << 4;
    const uint64_t seed = a.tin.seed_dev ? *a.tin.seed_dev : 0ULL;
    const uint32_t hseed = dropout_seed_hash(seed, a.tin.salt);
    const bool scaled = a.tin.escale != nullptr || a.tin.drop_p > 0.0f;
    const bool plain = a.tin.act == PCFD_ACT_NONE && !scaled;
    uint8_t* sslab = stg + slab * 4096;                 // phase 1 target (also the TMA-store buffer of warps 0..7)
    uint8_t* srow = sslab + lane * 128;
    uint32_t tl = 0;
    bool pending = false;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
      const int rt = t / a.k_passes, kp = t - rt * a.k_passes;
      const int64_t row0 = (int64_t)rt * POINTS;
      const uint32_t buf = tl & 1, aph = (tl >> 1) & 1;
      const int kk0 = kp * NT;
      int ncb = (a.k - kk0 + 31) / 32;
      if (ncb > NT / 32) ncb = NT / 32;
      const uint32_t tcol = tmem_base + ((uint32_t)(32 * q) << 16) + buf * (2 * NT) + u * NT;
      if (plain) {
        // ---- no input transform: staging tile -> TMA store, warps independent (the first 8 warps do it)
        tc::bounded_wait(&acc_full[buf], aph);
        tc::tc_fence_after();
        if (slab < SLABS && h == 0) {
          const int pg = slab / CJ, c = slab - pg * CJ;
#pragma unroll 1
          for (int cb = 0; cb < ncb; ++cb) {
            uint32_t v[32];
            tmem_ld32_nowait(tcol + cb * 32, v);
            tmem_ld_wait();
            if (pending) {
              if (lane == 0) tma_store_wait_read<0>();
              __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<uint4*>(srow + (((uint32_t)i << 4) ^ x7)) = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&tmO, sslab, kk0 + cb * 32, (int)(row0 + pg * 32), c);
              tma_store_commit();
            }
            pending = true;
          }
        }
        tc::tc_fence_before();
        mbar_arrive(&acc_free[buf]);
      } else {
        // ---- reverse activation jet through the staging tile
        const int64_t last_row = (row0 + POINTS <= a.rows ? row0 + POINTS : a.rows) - 1;
        const int64_t g_first = geom_of(row0, a.rows_per_geom);
        const bool uniform = g_first == geom_of(last_row, a.rows_per_geom);
        // pre-activation jets of this thread's items, loaded one block ahead
        float4 z[ITEMS][CJ];
        auto load_z = [&](int cb, float4 (&zz)[ITEMS][CJ]) {
          const int col = kk0 + cb * 32 + j * 4;
#pragma unroll
          for (int i = 0; i < ITEMS; ++i) {
            const int pgi = pgsel + 2 * i;
            const int64_t row = row0 + pgi * 32 + r;
#pragma unroll
            for (int c = 0; c < CJ; ++c) {
              zz[i][c] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (pgi < PGS && cb < ncb && col < a.k && row < a.rows)
                zz[i][c] = __ldg(reinterpret_cast<const float4*>(a.zin + c * a.zin_ps + row * a.ldzin + col));
            }
          }
        };
        load_z(0, z);
        tc::bounded_wait(&acc_full[buf], aph);
        tc::tc_fence_after();
#pragma unroll 1
        for (int cb = 0; cb < ncb; ++cb) {
          const int col = kk0 + cb * 32 + j * 4;
          const bool cvalid = col < a.k;
          // phase 1: accumulators -> staging (lane = row of the slab, this warp's 16 columns)
          if (slab < SLABS) {
            float v[16];
            tc::tmem_ld16(tcol + cb * 32 + h * 16, v);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<float4*>(srow + (((uint32_t)(4 * h + i) << 4) ^ x7)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          if (cb == ncb - 1) {
            tc::tc_fence_before();
            mbar_arrive(&acc_free[buf]);                 // TMEM buffer drained: the next tile's MMAs may start
          }
          float4 zn[ITEMS][CJ];
          load_z(cb + 1, zn);
          named_bar_sync(1, DX_EPI_THREADS);
          // phase 2: all channels of (row, 4 columns) per thread
          float gsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < ITEMS; ++i) {
            const int pgi = pgsel + 2 * i;
            const int64_t row = row0 + pgi * 32 + r;
            if (!(pgi < PGS && cvalid && row < a.rows) || (a.dbg & 16)) continue;
            float gq[CJ][4];
#pragma unroll
            for (int c = 0; c < CJ; ++c) {
              const float4 x = *reinterpret_cast<const float4*>(stg + (pgi * CJ + c) * 4096 + swz<128>(r, j));
              gq[c][0] = x.x; gq[c][1] = x.y; gq[c][2] = x.z; gq[c][3] = x.w;
            }
            if (col < a.tin.act_cols && !(a.dbg & 8)) {
              const int64_t geom = uniform ? g_first : geom_of(row, a.rows_per_geom);
              float ge[4] = {0.f, 0.f, 0.f, 0.f};
              reverse_dispatch<CJ>(gq, z[i], a.tin, scaled, hseed, row, geom, col, a.tin.act_cols - col, ge);
              if (a.gescale != nullptr) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (uniform) gsum[e] += ge[e];
                  else if (ge[e] != 0.0f) atomicAdd(a.gescale + geom * a.ldgescale + col + e, ge[e]);
                }
              }
            }
            if (!(a.dbg & 32)) {
#pragma unroll
              for (int c = 0; c < CJ; ++c)
                *reinterpret_cast<float4*>(a.gzin + c * a.gzin_ps + row * a.ldgzin + col) =
                    make_float4(gq[c][0], gq[c][1], gq[c][2], gq[c][3]);
            }
          }
          if (a.gescale != nullptr && uniform) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float v = gsum[e];
              v += __shfl_xor_sync(0xffffffffu, v, 8);
              v += __shfl_xor_sync(0xffffffffu, v, 16);
              if (lane < 8 && v != 0.0f) atomicAdd(&ge_acc[cb * 32 + j * 4 + e], v);
            }
          }
#pragma unroll
          for (int i = 0; i < ITEMS; ++i)
#pragma unroll
            for (int c = 0; c < CJ; ++c) z[i][c] = zn[i][c];
          named_bar_sync(2, DX_EPI_THREADS);             // staging tile free again, ge_acc complete
        }
        if (a.gescale != nullptr && uniform && et < NT) {
          const float v = ge_acc[et];
          if (kk0 + et < a.k && v != 0.0f) atomicAdd(a.gescale + g_first * a.ldgescale + kk0 + et, v);
          ge_acc[et] = 0.0f;
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == DX_W_MMA) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int CJ, int NT>
static int launch_dx(const float* gzout, int64_t gzout_ps, int ldgzout, const float* w, int ldw, DxArgs a, cudaStream_t st) {
  constexpr int PGS = 8 / CJ;
  constexpr int POINTS = 32 * PGS;
  constexpr int SMEM = STAGES * (2 * A_BYTES + 2 * NT * 64) + 8 * 4096 + 1024;
  CUtensorMap tmG, tmW, tmO;
  {
    const uint64_t dims[3] = {(uint64_t)a.n, (uint64_t)a.rows, (uint64_t)CJ};
    const uint64_t str[2] = {(uint64_t)ldgzout * 4, (uint64_t)gzout_ps * 4};
    const uint32_t box[3] = {BK, CJ == 1 ? 256u : 32u, (uint32_t)CJ};
    if (!make_tmap(&tmG, gzout, 3, dims, str, box, 64)) return PCFD_ERR_ARG;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.k, (uint64_t)a.n};
    const uint64_t str[1] = {(uint64_t)ldw * 4};
    const uint32_t box[2] = {32, BK};
    if (!make_tmap(&tmW, w, 2, dims, str, box, SW128_ATOM32)) return PCFD_ERR_ARG;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a.k, (uint64_t)a.rows, (uint64_t)CJ};
    const uint64_t str[2] = {(uint64_t)a.ldgzin * 4, (uint64_t)a.gzin_ps * 4};
    const uint32_t box[3] = {32, 32, 1};
    if (!make_tmap(&tmO, a.gzin, 3, dims, str, box, 128)) return PCFD_ERR_ARG;
  }
  a.row_tiles = (int)((a.rows + POINTS - 1) / POINTS);
  a.k_passes = (a.k + NT - 1) / NT;
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("PCFD_WS_DEBUG"); dbg = e ? atoi(e) : 0; }
  a.dbg = dbg;
  {
    const cudaError_t e = ensure_dyn_smem<ws_dx_kernel<CJ, NT>>(SMEM);
    if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;
  }
  const int total = a.row_tiles * a.k_passes;
  const int grid = balanced_grid(total);
  const cudaError_t le = launch_pdl(ws_dx_kernel<CJ, NT>, dim3(grid), dim3(DX_THREADS), (size_t)SMEM, st, tmG, tmW, tmO, a);
  if (le != cudaSuccess) return PCFD_ERR_CUDA + (int)le;
  return PCFD_OK;
}

}  // namespace ws
}  // namespace pcfd

using namespace pcfd;

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int pcfd_ws_supported_dx(const float* gzout, int64_t gzout_ps, int32_t ldgzout, const float* w, int32_t ldw,
                                    const float* zin, int64_t zin_ps, int32_t ldzin, const float* gzin, int64_t gzin_ps,
                                    int32_t ldgzin, int32_t cj, int64_t rows, int32_t k, int32_t n) {
  if (!valid_cj(cj) || rows < 256 || k < 16 || n < 8) return 0;
  if (rows >= (int64_t)1 << 31) return 0;
  if (!al16(gzout) || !al16(w) || !al16(zin) || !al16(gzin)) return 0;
  if (ldgzout % 4 || ldw % 4 || ldzin % 4 || ldgzin % 4) return 0;
  if (cj > 1 && (gzout_ps % 4 || zin_ps % 4 || gzin_ps % 4)) return 0;
  return ws::encode_fn() != nullptr;
}

extern "C" int pcfd_ws_jet_linear_bwd_dx(const float* gzout, int64_t gzout_ps, int32_t ldgzout, const float* w,
                                         int32_t ldw, const float* zin, int64_t zin_ps, int32_t ldzin,
                                         const pcfd_intrans_t* tin, float* gzin, int64_t gzin_ps, int32_t ldgzin,
                                         float* gescale, int32_t ldgescale, int32_t cj, int64_t rows,
                                         int64_t rows_per_geom, int32_t k, int32_t n, void* stream) {
  if (cj == 1) { gzout_ps = (int64_t)rows * ldgzout; zin_ps = (int64_t)rows * ldzin; gzin_ps = (int64_t)rows * ldgzin; }
  ws::DxArgs a{zin, zin_ps, ldzin, gzin, gzin_ps, ldgzin, gescale, ldgescale, rows, rows_per_geom, k, n,
               make_intrans(tin, k), 0, 0, 0};
  cudaStream_t st = (cudaStream_t)stream;
  // few row tiles (per-geometry layers): 64-column passes double the number of CTAs; otherwise 128-column tiles
  const int pts = 32 * (8 / cj);
  const int64_t items128 = ((rows + pts - 1) / pts) * ((k + 127) / 128);
  const bool narrow = k <= 64 || items128 * 2 <= ws::num_sms();
#define PCFD_WS_DX(CJ_)                                                             \
  return narrow ? ws::launch_dx<CJ_, 64>(gzout, gzout_ps, ldgzout, w, ldw, a, st)  \
                 : ws::launch_dx<CJ_, 128>(gzout, gzout_ps, ldgzout, w, ldw, a, st);
  switch (cj) {
    case 1: PCFD_WS_DX(1)
    case 3: PCFD_WS_DX(3)
    case 4: PCFD_WS_DX(4)
    case 5: PCFD_WS_DX(5)
    case 7: PCFD_WS_DX(7)
  }
#undef PCFD_WS_DX
  return PCFD_ERR_ARG;
}
