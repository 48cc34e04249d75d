// Engine 2, backward of a jet layer to its weights: warp-specialised tcgen05 kernel.
//
//   partial[split][n][k] = sum over (channel c, row in split)  gzout[c][row][n] * T(zin)[c][row][k]
//
// The contraction runs over the ROWS (points x channels), which is the slow index of both operands in
// memory, so both are MN-major UMMA operands: TMA brings [R rows x 32 columns x CJ channels] boxes in
// the 32-byte-atom 128B swizzle (ws_common.cuh) and the tensor core transposes them -- neither
// gzout^T nor T(zin)^T is ever materialised.  A stage holds E = CJ*R contraction entries (all channels
// of R points, so the activation jet can be applied to the staged tile in place) of
//   A = gzout   columns [n0, n0 + MT*128)      (MT UMMA M-tiles of 128 layer outputs)
//   B = T(zin)  columns [k0, k0 + NTL*NT)      (NTL UMMA N-tiles of NT layer inputs)
// and the MT x NTL accumulator tiles (<= 512 TMEM columns) stay resident for the whole row range of the
// CTA.  grid = (passes over the (n, k) plane, row splits); the splits' partial products are reduced in a
// fixed order by pcfd_dw_finish (deterministic), which also forms the bias / per-geometry gradients.
//
//   warps 0-15  1-4 transform groups (plan_dw: few large groups for short rings); group g transforms ring iterations
//               g, g + groups, ...: activation jet /
//               dropout / branch scaling of the zin tile in place (= TF32 "hi" operand, the tensor core
//               reads the top 19 bits) and the exact remainders lo = x - trunc_tf32(x) of both operands;
//               after the main loop the same warps drain the accumulators to the partial buffer
//   warp 16     TMA producer, warp 17 MMA issuer (D += Ahi*Bhi + Alo*Bhi + Ahi*Blo)
//   warp 18     column sums of the value plane of the staged gzout tile (the bias gradient), kept in registers over
//               the whole row range and written as one partial row per split -- saves a second pass over gzout
#include <cstdlib>

#include "common.cuh"
#include "ws_common.cuh"

namespace pcfd {
namespace ws {

#ifndef PCFD_DW_MAX_GROUPS
#define PCFD_DW_MAX_GROUPS 4
#endif
constexpr int DW_XF_THREADS = 512;      // transform threads (16 warps), split into a.groups groups at run time
constexpr int DW_W_TMA = DW_XF_THREADS / 32, DW_W_MMA = DW_W_TMA + 1, DW_W_SUM = DW_W_TMA + 2;
constexpr int DW_THREADS = (DW_W_SUM + 1) * 32;
constexpr int DW_MAX_STAGES = 8;
constexpr int DW_SMEM_MAX = 232448 - 2048;      // dynamic shared memory we may ask for (227 KB less static + alignment slack)

struct DwArgs {
  float* partial;
  int64_t rows, rows_per_geom, rows_per_split;
  int k, n;
  InTrans tin;
  int mt, ntl;                  // accumulator tiles of one pass: mt x 128 outputs, ntl x NT inputs
  int passes_k;                 // passes along k (pass index = pn * passes_k + pk)
  int stages, groups;           // ring depth; transform groups in use (groups divides stages, see plan_dw)
  int vec_out;                  // k % 4 == 0: 16-byte stores into the partial buffer
  float* colsum;                // optional [splits][n]: column sums of plane 0 of gzout over the split's rows
};

template <int CJ, int R, int NT>
__global__ void __launch_bounds__(DW_THREADS, 1) ws_dw_kernel(const __grid_constant__ CUtensorMap tmG,
                                                              const __grid_constant__ CUtensorMap tmZ, DwArgs a) {
  constexpr int E = CJ * R;                       // contraction entries per stage
  constexpr int BLK = E * 128;                    // one 32-column block of a stage
  static_assert(E % 8 == 0 && R % 4 == 0 && (R & (R - 1)) == 0, "stage must hold whole MMA K steps");
  static_assert(BLK % 1024 == 0, "blocks must keep the swizzle phase");
  constexpr int NB = NT / 32;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t raw_full[DW_MAX_STAGES], ops_ready[DW_MAX_STAGES], stage_free[DW_MAX_STAGES];
  __shared__ __align__(8) uint64_t acc_full;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = uniform_warp_id(), lane = tid & 31;
  const int STG = a.stages;
  const int pn = (int)blockIdx.x / a.passes_k, pk = (int)blockIdx.x - pn * a.passes_k;
  const int n0 = pn * a.mt * 128, k0 = pk * a.ntl * NT;
  const int ncols = min(a.n - n0, a.mt * 128), kcols = min(a.k - k0, a.ntl * NT);
  const int ab = (ncols + 31) >> 5, bb = (kcols + 31) >> 5;      // 32-column blocks actually loaded
  const int AB = a.mt * 4, BB = a.ntl * NB;                        // blocks the stage layout reserves
  const uint32_t A_LO = (uint32_t)AB * BLK, B_HI = 2u * AB * BLK, B_LO = B_HI + (uint32_t)BB * BLK;
  const uint32_t STAGE_BYTES = 2u * (AB + BB) * BLK;
  const int64_t r_begin = (int64_t)blockIdx.y * a.rows_per_split;
  const int64_t r_end = min(a.rows, r_begin + a.rows_per_split);
  const int nsteps = (int)((r_end - r_begin + R - 1) / R);

  if (tid == 0) {
    for (int s = 0; s < STG; ++s) {
      tc::mbar_init(&raw_full[s], 1);
      tc::mbar_init(&ops_ready[s], (16 / a.groups) * 32 + (a.colsum != nullptr ? 32 : 0));
      tc::mbar_init(&stage_free[s], 1);
    }
    tc::mbar_init(&acc_full, 1);
    tc::fence_mbar_init();
  }
  if (warp == DW_W_MMA) tc::tmem_alloc(&tmem_base_s, 512);
  if (warp == DW_W_TMA && lane == 0) { prefetch_tmap(&tmG); prefetch_tmap(&tmZ); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  griddep_wait();                  // everything above overlaps the tail of the previous kernel of the stream
  griddep_launch_dependents();     // one resident wave: the next kernel may take SMs as they free up

  if (warp == DW_W_TMA) {
    // ================================ TMA producer ================================
    const uint32_t tx = (uint32_t)(ab + bb) * BLK;
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nsteps; ++it) {
      tc::bounded_wait(&stage_free[s], ph ^ 1);
      uint8_t* st = smem + (size_t)s * STAGE_BYTES;
      const int row = (int)r_begin + it * R;
      if (elect_one()) {
        mbar_expect_tx(&raw_full[s], tx);
        for (int b = 0; b < ab; ++b) tma_load_3d(st + b * BLK, &tmG, n0 + b * 32, row, 0, &raw_full[s]);
        for (int b = 0; b < bb; ++b) tma_load_3d(st + B_HI + b * BLK, &tmZ, k0 + b * 32, row, 0, &raw_full[s]);
      }
      __syncwarp();
      if (++s == STG) { s = 0; ph ^= 1; }
    }
  } else if (warp == DW_W_MMA) {
    // ================================ MMA issuer ================================
    constexpr uint32_t IDESC = tc::make_idesc_tf32(128, NT, true, true);       // both operands MN-major
    const uint64_t dbase = desc_mnmajor(tc::smem_u32(smem), BLK, 512);        // + (byte offset >> 4) in the address field
    const uint32_t a_lo = A_LO >> 4, b_hi = B_HI >> 4, b_lo = B_LO >> 4;
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nsteps; ++it) {
      tc::bounded_wait(&ops_ready[s], ph);
      tc::tc_fence_after();
      const uint64_t ds = dbase + (uint64_t)(((uint32_t)s * STAGE_BYTES) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < E / 8; ++ks) {
          for (int mt = 0; mt < a.mt; ++mt) {
            const uint64_t da_hi = ds + (uint32_t)((mt * 4 * BLK + ks * 1024) >> 4);
            const uint64_t da_lo = da_hi + a_lo;
            for (int nl = 0; nl < a.ntl; ++nl) {
              const uint64_t db_hi = ds + b_hi + (uint32_t)((nl * NB * BLK + ks * 1024) >> 4);
              const uint64_t db_lo = db_hi + (b_lo - b_hi);
              const uint32_t d = tmem_base + (uint32_t)(mt * a.ntl + nl) * NT;
              tc::mma_tf32(d, da_hi, db_hi, IDESC, (it > 0 || ks > 0) ? 1u : 0u);
              tc::mma_tf32(d, da_lo, db_hi, IDESC, 1u);
              tc::mma_tf32(d, da_hi, db_lo, IDESC, 1u);
            }
          }
        }
        tc::mma_commit(&stage_free[s]);
      }
      __syncwarp();
      if (++s == STG) { s = 0; ph ^= 1; }
    }
    if (elect_one()) tc::mma_commit(&acc_full);
    __syncwarp();
  } else if (warp == DW_W_SUM) {
    // ================================ bias column sums ================================
    // lane = (block within a group of 4, logical 16-byte chunk); rows of channel 0 only (contraction entries < R)
    if (a.colsum != nullptr) {
      const int jl = lane & 7, bl = lane >> 3;
      float4 acc[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < nsteps; ++it) {
        const uint8_t* st = smem + (size_t)s * STAGE_BYTES;
        tc::bounded_wait(&raw_full[s], ph);
        if (pk == 0) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int b = 4 * t + bl;
            if (b < ab) {
#pragma unroll
              for (int e = 0; e < R; ++e) {
                const uint32_t off = (uint32_t)e * 128u + (((((uint32_t)jl >> 1) ^ ((uint32_t)e & 3u)) << 5) | (((uint32_t)jl & 1u) << 4));
                const float4 x = *reinterpret_cast<const float4*>(st + b * BLK + off);
                acc[t].x += x.x; acc[t].y += x.y; acc[t].z += x.z; acc[t].w += x.w;
              }
            }
          }
        }
        mbar_arrive(&ops_ready[s]);
        if (++s == STG) { s = 0; ph ^= 1; }
      }
      if (pk == 0) {
        float* dst = a.colsum + (int64_t)blockIdx.y * a.n;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int col = n0 + (4 * t + bl) * 32 + jl * 4;
          if (4 * t + bl < ab) {
            if (col < a.n) dst[col] = acc[t].x;
            if (col + 1 < a.n) dst[col + 1] = acc[t].y;
            if (col + 2 < a.n) dst[col + 2] = acc[t].z;
            if (col + 3 < a.n) dst[col + 3] = acc[t].w;
          }
        }
      }
    }
  } else {
    // ================================ transform ================================
    const int gs = (16 / a.groups) * 32;              // threads of one transform group (whole warps)
    const int g = warp / (gs >> 5);
    const int tt = tid - g * gs;
    const uint64_t seed = a.tin.seed_dev ? *a.tin.seed_dev : 0ULL;
    const uint32_t hseed = dropout_seed_hash(seed, a.tin.salt);
    const bool scaled = a.tin.escale != nullptr || a.tin.drop_p > 0.0f;
    const bool plain = a.tin.act == PCFD_ACT_NONE && !scaled;
    const int a_chunks = ab * (BLK / 16);
    const int b_chunks = bb * (BLK / 16);
    const int b_items = bb * R * 8;                 // (block, row, 16-byte chunk) positions, all channels each
    int s = g;                                      // g < groups <= stages
    uint32_t ph = 0;
    for (int it = g; it < (g < a.groups ? nsteps : 0); it += a.groups) {
      uint8_t* st = smem + (size_t)s * STAGE_BYTES;
      const int64_t row0 = r_begin + (int64_t)it * R;
      tc::bounded_wait(&raw_full[s], ph);
      // ---- A: remainder tile of gzout (same swizzled position in the lo tile)
      for (int i0 = tt; i0 < a_chunks; i0 += 4 * gs) {
        float4 x[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (i0 + q * gs < a_chunks) x[q] = *reinterpret_cast<const float4*>(st + (size_t)(i0 + q * gs) * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (i0 + q * gs < a_chunks)
            *reinterpret_cast<float4*>(st + A_LO + (size_t)(i0 + q * gs) * 16) =
                make_float4(x[q].x - trunc_tf32(x[q].x), x[q].y - trunc_tf32(x[q].y), x[q].z - trunc_tf32(x[q].z),
                            x[q].w - trunc_tf32(x[q].w));
      }
      uint8_t* bt = st + B_HI;
      if (plain) {
        for (int i0 = tt; i0 < b_chunks; i0 += 4 * gs) {
          float4 x[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (i0 + q * gs < b_chunks) x[q] = *reinterpret_cast<const float4*>(bt + (size_t)(i0 + q * gs) * 16);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (i0 + q * gs < b_chunks)
              *reinterpret_cast<float4*>(bt + (B_LO - B_HI) + (size_t)(i0 + q * gs) * 16) =
                  make_float4(x[q].x - trunc_tf32(x[q].x), x[q].y - trunc_tf32(x[q].y), x[q].z - trunc_tf32(x[q].z),
                              x[q].w - trunc_tf32(x[q].w));
        }
      } else {
        // ---- B: activation jet in place + remainder tile
        for (int idx = tt; idx < b_items; idx += gs) {
          const int j = idx & 7, r = (idx >> 3) & (R - 1), blk = idx / (8 * R);
          uint8_t* base = bt + blk * BLK + r * 128 + ((((uint32_t)(j >> 1) ^ (uint32_t)(r & 3)) << 5) | ((uint32_t)(j & 1) << 4));
          float v[CJ][4];
#pragma unroll
          for (int c = 0; c < CJ; ++c) {
            const float4 x = *reinterpret_cast<const float4*>(base + c * (R * 128));
            v[c][0] = x.x; v[c][1] = x.y; v[c][2] = x.z; v[c][3] = x.w;
          }
          const int64_t row = row0 + r;
          const int col0 = k0 + blk * 32 + j * 4;
          if (row < a.rows && col0 < a.tin.act_cols) {
            const int64_t geom = a.tin.escale != nullptr ? geom_of(row, a.rows_per_geom) : 0;
            transform_dispatch<CJ>(v, a.tin, scaled, hseed, row, geom, col0, a.tin.act_cols - col0);
#pragma unroll
            for (int c = 0; c < CJ; ++c)
              *reinterpret_cast<float4*>(base + c * (R * 128)) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
          }
#pragma unroll
          for (int c = 0; c < CJ; ++c)
            *reinterpret_cast<float4*>(base + (B_LO - B_HI) + c * (R * 128)) =
                make_float4(v[c][0] - trunc_tf32(v[c][0]), v[c][1] - trunc_tf32(v[c][1]), v[c][2] - trunc_tf32(v[c][2]),
                            v[c][3] - trunc_tf32(v[c][3]));
        }
      }
      tc::fence_proxy_async();
      mbar_arrive(&ops_ready[s]);
      s += a.groups;
      if (s >= STG) { s -= STG; ph ^= 1; }
    }

    // ================================ epilogue: accumulators -> partial[split] ================================
    const int q = warp & 3;                           // TMEM lane quarter this warp may read
    const int cgrp = warp >> 2;                       // column blocks cgrp, cgrp + 4, ...
    const int blocks_per_mt = a.ntl * NB;
    const int total_cb = a.mt * blocks_per_mt;
    float* dst = a.partial + (int64_t)blockIdx.y * a.n * a.k;
    if (nsteps > 0) {
      tc::bounded_wait(&acc_full, 0);
      tc::tc_fence_after();
    }
#pragma unroll 1
    for (int cbi = cgrp; cbi < total_cb; cbi += 4) {
      const int mt = cbi / blocks_per_mt;
      const int kl = (cbi - mt * blocks_per_mt) * 32;
      const int nn = n0 + mt * 128 + 32 * q + lane;
      const int kk0 = k0 + kl;
      if (kl >= kcols || mt * 128 >= ncols) continue;           // warp-uniform
      uint32_t v[32];
      if (nsteps > 0) {
        tmem_ld32_nowait(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)cbi * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      if (nn < a.n) {
        float* p = dst + (int64_t)nn * a.k + kk0;
        if (a.vec_out && kk0 + 32 <= a.k) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<uint4*>(p + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kk0 + i < a.k) p[i] = __uint_as_float(v[i]);
        }
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == DW_W_MMA) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace ws
}  // namespace pcfd

#include "ws_dw1.cuh"

namespace pcfd {
namespace ws {

// ---- host: pass / split plan -------------------------------------------------------------------
struct DwPlan {
  int kind;                           // 0: ws_dw_kernel (operands in shared memory), 1: ws_dw1_kernel (cj = 1, gzout in TMEM)
  int r, e, nt, mt, ntl, passes_n, passes_k, stages, groups, splits;
  int64_t rows_per_split;
  int64_t chunks, rows_per_chunk;     // column-sum scratch of pcfd_dw_finish
};

// rows per stage: the small value keeps wide tiles within shared memory, the large one amortises the
// per-stage barrier round trips of narrow layers
static inline int dw_rows_small(int cj) { return cj == 1 ? 16 : (cj == 4 ? 4 : 8); }
// cj = 4: 16 rows never left room for 3 stages (96 KB per stage on the narrowest layer); with 8 rows the narrow layers
// (3->64, 64->64, 128->4) run half as many ring iterations per CTA -- their time was the per-iteration latency chain
// (ncu: 45 % of the samples in the transform warps' wait for the TMA barrier at 4 rows per stage)
static inline int dw_rows_large(int cj) { return cj == 1 ? 64 : (cj == 4 ? 8 : (cj == 3 ? 16 : 8)); }

// a waiter may only be one phase away from its barrier: the transform group of iteration `it` must also be
// the group of iteration `it - stages`, so the number of groups in use divides the ring depth
static inline void dw_ring(int fit, int* stages, int* groups) {
  if (fit >= 8) { *stages = 8; *groups = 4; }
  else if (fit >= 6) { *stages = 6; *groups = 3; }
  else if (fit >= 4) { *stages = 4; *groups = 4; }
  else { *stages = fit; *groups = fit; }
  // With a short ring the time a stage spends in the transform is on the critical path: fewer, larger groups (measured:
  // 1 group of 16 warps at 3 stages 154 -> 141 us on the 384->128 layer); with 6-8 small stages 3-4 groups are better.
  if (*stages <= 4) *groups = (*stages % 2 == 0) ? 2 : 1;
  while (*groups > PCFD_DW_MAX_GROUPS || (*groups > 0 && *stages % *groups != 0)) --*groups;
}

// row splits of a pass plan: one CTA per (pass, split), at least a few ring iterations each
static void plan_splits(DwPlan& p, int cj, int64_t rows, int64_t rows_per_geom) {
  p.e = cj * p.r;
  const int passes = p.passes_n * p.passes_k;
  int64_t splits = num_sms() / passes;
  if (splits < 1) splits = 1;
  // at least a few ring iterations per CTA.  Measured (PCFD_DW_MIN_ITERS = 2..64 on the abc layers): a CTA's pipeline is
  // latency-bound, so more, shorter splits win over fewer partials -- 8 and below are equal, 32 costs 20-50 % on the
  // small layers
  static int min_iters = -1;
  if (min_iters < 0) { const char* e = getenv("PCFD_DW_MIN_ITERS"); min_iters = e ? atoi(e) : 8; }
  const int64_t min_rows = (int64_t)min_iters * p.r;
  const int64_t max_splits = (rows + min_rows - 1) / min_rows;
  if (splits > max_splits) splits = max_splits;
  int64_t rps = (rows + splits - 1) / splits;
  rps = (rps + p.r - 1) / p.r * p.r;
  p.rows_per_split = rps;
  p.splits = (int)((rows + rps - 1) / rps);
  p.rows_per_chunk = rows_per_geom > 0 ? rows_per_geom : 2048;
  p.chunks = (rows + p.rows_per_chunk - 1) / p.rows_per_chunk;
}


// ring of the TMEM-operand kernel: groups are 4 warps (one per TMEM lane quarter) and divide the ring depth
static inline void dw1_ring(int fit, int* stages, int* groups) {
  if (fit >= 8) { *stages = 8; *groups = 4; }
  else if (fit >= 6) { *stages = 6; *groups = 3; }
  else if (fit >= 4) { *stages = 4; *groups = 4; }
  else if (fit >= 2) { *stages = fit; *groups = fit; }
  else { *stages = 0; *groups = 0; }
}

// PCFD_DW1=0 keeps value-only layers on ws_dw_kernel; PCFD_DW1_R = 16 / 32 forces the rows per stage
static int dw1_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("PCFD_DW1"); on = e ? atoi(e) : 1; }
  return on;
}

static DwPlan plan_dw1(int64_t rows, int64_t rows_per_geom, int k, int n) {
  DwPlan p{};
  p.kind = 1;
  p.nt = k <= 64 ? 64 : 128;
  const int mt_all = (n + 127) / 128, ntl_all = (k + p.nt - 1) / p.nt;
  static int force_r = -1;
  if (force_r < 0) { const char* e = getenv("PCFD_DW1_R"); force_r = e ? atoi(e) : 0; }
  double best = 1e300;
  for (int r = 16; r <= 32; r *= 2) {
    if (force_r && r != force_r) continue;
    for (int mt = 1; mt <= mt_all && mt <= 4; ++mt) {
      for (int ntl = 1; ntl <= ntl_all; ++ntl) {
        const int acc = mt * ntl * p.nt;                       // accumulator columns; the rest of TMEM holds operand slots
        if (acc + 2 * mt * 2 * r > 512) continue;
        const int stage_bytes = (mt * 4 + 2 * ntl * p.nt / 32) * (r * 128);
        int fit = DW_SMEM_MAX / stage_bytes;
        if ((512 - acc) / (mt * 2 * r) < fit) fit = (512 - acc) / (mt * 2 * r);
        int stages = 0, groups = 0;
        dw1_ring(fit, &stages, &groups);
        if (stages < 2) continue;
        const int pn = (mt_all + mt - 1) / mt, pk = (ntl_all + ntl - 1) / ntl;
        double cost = (double)n * pk + (double)k * pn;
        if (stages < 3) cost *= 1.5;
        if (r == 32 && stages >= 4) cost *= 0.999;             // same traffic: prefer the longer stage
        if (cost < best - 1e-9) {
          best = cost; p.mt = mt; p.ntl = ntl; p.stages = stages; p.groups = groups; p.passes_n = pn; p.passes_k = pk; p.r = r;
        }
      }
    }
  }
  if (p.stages >= 2) plan_splits(p, 1, rows, rows_per_geom);
  return p;
}

static DwPlan plan_dw(int cj, int64_t rows, int64_t rows_per_geom, int k, int n) {
  if (cj == 1 && dw1_enabled()) {
    const DwPlan p1 = plan_dw1(rows, rows_per_geom, k, n);
    if (p1.stages >= 2) return p1;
  }
  DwPlan p;
  p.kind = 0;
  p.nt = k <= 64 ? 64 : 128;
  const int mt_all = (n + 127) / 128, ntl_all = (k + p.nt - 1) / p.nt;
  static int force_r = -1;
  if (force_r < 0) { const char* e = getenv("PCFD_DW_R"); force_r = e ? atoi(e) : 0; }
  double best = 1e300;
  p.mt = 1; p.ntl = 1; p.stages = 0; p.groups = 0; p.r = dw_rows_small(cj);
  for (int mt = 1; mt <= mt_all && mt <= 4; ++mt) {
    for (int ntl = 1; ntl <= ntl_all && mt * ntl * p.nt <= 512; ++ntl) {
      for (int big = 0; big < 2; ++big) {
        const int r = big ? dw_rows_large(cj) : dw_rows_small(cj);
        if (big && r == dw_rows_small(cj)) continue;
        if (force_r == 1 && big) continue;
        const int stage_bytes = 2 * (mt * 4 + ntl * p.nt / 32) * (cj * r * 128);
        int stages = 0, groups = 0;
        dw_ring(DW_SMEM_MAX / stage_bytes, &stages, &groups);
        if (stages < 2) continue;
        if (big && stages < 3 && force_r != 2) continue;
        const int pn = (mt_all + mt - 1) / mt, pk = (ntl_all + ntl - 1) / ntl;
        // HBM traffic of the operands (each pass along k re-reads gzout, each pass along n re-reads zin);
        // fewer than 3 stages cannot cover the TMA -> transform -> MMA latency chain
        double cost = (double)n * pk + (double)k * pn;
        if (stages < 3) cost *= 1.5;
        if (big) cost *= 0.999;                                  // same traffic: prefer the longer stage
        if (cost < best - 1e-9) {
          best = cost; p.mt = mt; p.ntl = ntl; p.stages = stages; p.groups = groups; p.passes_n = pn; p.passes_k = pk; p.r = r;
        }
      }
    }
  }
  plan_splits(p, cj, rows, rows_per_geom);
  return p;
}

template <int CJ, int R, int NT>
static int launch_dw(const float* gzout, int64_t gzout_ps, int ldgzout, const float* zin, int64_t zin_ps, int ldzin,
                     DwArgs a, const DwPlan& p, cudaStream_t st) {
  CUtensorMap tmG, tmZ;
  {
    const uint64_t dims[3] = {(uint64_t)a.n, (uint64_t)a.rows, (uint64_t)CJ};
    const uint64_t str[2] = {(uint64_t)ldgzout * 4, (uint64_t)gzout_ps * 4};
    const uint32_t box[3] = {32, (uint32_t)R, (uint32_t)CJ};
    if (!make_tmap(&tmG, gzout, 3, dims, str, box, SW128_ATOM32)) return PCFD_ERR_ARG;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a.k, (uint64_t)a.rows, (uint64_t)CJ};
    const uint64_t str[2] = {(uint64_t)ldzin * 4, (uint64_t)zin_ps * 4};
    const uint32_t box[3] = {32, (uint32_t)R, (uint32_t)CJ};
    if (!make_tmap(&tmZ, zin, 3, dims, str, box, SW128_ATOM32)) return PCFD_ERR_ARG;
  }
  const int stage_bytes = 2 * (p.mt * 4 + p.ntl * NT / 32) * (CJ * R * 128);
  const int smem = p.stages * stage_bytes + 1024;
  if (smem > DW_SMEM_MAX + 1024) return PCFD_ERR_ARG;
  {
    const cudaError_t e = ensure_dyn_smem<ws_dw_kernel<CJ, R, NT>>(DW_SMEM_MAX + 1024);
    if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;
  }
  dim3 grid((unsigned)(p.passes_n * p.passes_k), (unsigned)p.splits);
  const cudaError_t le = launch_pdl(ws_dw_kernel<CJ, R, NT>, grid, dim3(DW_THREADS), (size_t)smem, st, tmG, tmZ, a);
  if (le != cudaSuccess) return PCFD_ERR_CUDA + (int)le;
  return PCFD_OK;
}


template <int R, int NT>
static int launch_dw1(const float* gzout, int ldgzout, const float* zin, int ldzin, DwArgs a, const DwPlan& p, cudaStream_t st) {
  CUtensorMap tmG, tmZ;
  {
    const uint64_t dims[2] = {(uint64_t)a.n, (uint64_t)a.rows};
    const uint64_t str[1] = {(uint64_t)ldgzout * 4};
    const uint32_t box[2] = {32, (uint32_t)R};
    if (!make_tmap(&tmG, gzout, 2, dims, str, box, SW128_ATOM32)) return PCFD_ERR_ARG;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.k, (uint64_t)a.rows};
    const uint64_t str[1] = {(uint64_t)ldzin * 4};
    const uint32_t box[2] = {32, (uint32_t)R};
    if (!make_tmap(&tmZ, zin, 2, dims, str, box, SW128_ATOM32)) return PCFD_ERR_ARG;
  }
  const int stage_bytes = (p.mt * 4 + 2 * p.ntl * NT / 32) * (R * 128);
  int smem = p.stages * stage_bytes;
  if (smem < 4 * 4 * 128 * 4) smem = 4 * 4 * 128 * 4;          // the column sums of the groups reuse the ring
  smem += 1024;
  if (smem > DW_SMEM_MAX + 1024) return PCFD_ERR_ARG;
  {
    const cudaError_t e = ensure_dyn_smem<ws_dw1_kernel<R, NT>>(DW_SMEM_MAX + 1024);
    if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;
  }
  dim3 grid((unsigned)(p.passes_n * p.passes_k), (unsigned)p.splits);
  const cudaError_t le = launch_pdl(ws_dw1_kernel<R, NT>, grid, dim3(DW1_THREADS), (size_t)smem, st, tmG, tmZ, a);
  if (le != cudaSuccess) return PCFD_ERR_CUDA + (int)le;
  return PCFD_OK;
}

}  // namespace ws
}  // namespace pcfd

using namespace pcfd;

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int pcfd_ws_supported_dw(const float* gzout, int64_t gzout_ps, int32_t ldgzout, const float* zin, int64_t zin_ps,
                                    int32_t ldzin, int32_t cj, int64_t rows, int32_t k, int32_t n) {
  if (!valid_cj(cj) || rows < 512 || (int64_t)k * n < 64) return 0;
  if (rows >= (int64_t)1 << 31) return 0;
  if (!al16(gzout) || !al16(zin)) return 0;
  if (ldgzout % 4 || ldzin % 4) return 0;
  if (cj > 1 && (gzout_ps % 4 || zin_ps % 4)) return 0;
  return ws::encode_fn() != nullptr;
}

extern "C" size_t pcfd_ws_dw_workspace_bytes(int32_t cj, int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n) {
  ws::DwPlan p = ws::plan_dw(cj, rows, rows_per_geom, k, n);
  return ((size_t)p.splits * n * k + (size_t)p.splits * n + (size_t)p.chunks * ((p.rows_per_chunk + 127) / 128) * n) *
             sizeof(float) + 256;
}

// writes partial[splits][n][k] at the start of `workspace` and, when `want_colsum`, the column sums of plane 0 of
// gzout per split right behind it ([splits][n]); returns the number of splits through *splits_out
extern "C" int pcfd_ws_jet_linear_bwd_dw_partials(const float* gzout, int64_t gzout_ps, int32_t ldgzout, const float* zin,
                                                  int64_t zin_ps, int32_t ldzin, const pcfd_intrans_t* tin, int32_t cj,
                                                  int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n,
                                                  void* workspace, int want_colsum, int* splits_out, void* stream) {
  ws::DwPlan p = ws::plan_dw(cj, rows, rows_per_geom, k, n);
  if (p.stages < 2) return PCFD_ERR_ARG;
  if (cj == 1) { gzout_ps = (int64_t)rows * ldgzout; zin_ps = (int64_t)rows * ldzin; }
  ws::DwArgs a{reinterpret_cast<float*>(workspace), rows, rows_per_geom, p.rows_per_split, k, n, make_intrans(tin, k),
               p.mt, p.ntl, p.passes_k, p.stages, p.groups, (k % 4 == 0 && al16(workspace)) ? 1 : 0,
               want_colsum ? reinterpret_cast<float*>(workspace) + (size_t)p.splits * n * k : nullptr};
  *splits_out = p.splits;
  cudaStream_t st = (cudaStream_t)stream;
  if (p.kind == 1) {
    if (p.r == 16)
      return p.nt == 64 ? ws::launch_dw1<16, 64>(gzout, ldgzout, zin, ldzin, a, p, st)
                        : ws::launch_dw1<16, 128>(gzout, ldgzout, zin, ldzin, a, p, st);
    return p.nt == 64 ? ws::launch_dw1<32, 64>(gzout, ldgzout, zin, ldzin, a, p, st)
                      : ws::launch_dw1<32, 128>(gzout, ldgzout, zin, ldzin, a, p, st);
  }
#define PCFD_WS_DW(CJ_, R_)                                                                                    \
  if (p.r == R_)                                                                                               \
    return p.nt == 64 ? ws::launch_dw<CJ_, R_, 64>(gzout, gzout_ps, ldgzout, zin, zin_ps, ldzin, a, p, st)     \
                      : ws::launch_dw<CJ_, R_, 128>(gzout, gzout_ps, ldgzout, zin, zin_ps, ldzin, a, p, st);
  switch (cj) {
    case 1: PCFD_WS_DW(1, 16) PCFD_WS_DW(1, 64) break;
    case 3: PCFD_WS_DW(3, 8) PCFD_WS_DW(3, 16) break;
    case 4: PCFD_WS_DW(4, 4) PCFD_WS_DW(4, 8) break;
    case 5: PCFD_WS_DW(5, 8) break;
    case 7: PCFD_WS_DW(7, 8) break;
  }
#undef PCFD_WS_DW
  return PCFD_ERR_ARG;
}
