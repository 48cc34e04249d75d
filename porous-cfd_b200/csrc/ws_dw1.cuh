// Engine 2, weight gradient of a VALUE-ONLY layer (cj = 1) with the gzout operand in tensor memory.  Included by ws_dw.cu
// (shares DwArgs / DwPlan and the partial-buffer convention of ws_dw_kernel).
//
//   partial[split][n][k] = sum over the split's rows of  gzout[row][n] * T(zin)[row][k]
//
// ws_dw_kernel keeps both MN-major operands and their TF32 remainders in shared memory; for one channel a stage is only
// R = 16 rows deep, and the hi/lo stores plus the tensor core's three reads of the gzout tile load the shared-memory pipe
// more than the MMAs load the tensor pipe.  Here the D rows (TMEM lanes) are the gzout COLUMNS, so the thread that owns
// lane n reads column n of the staged tile (32 lanes = one 128-byte row: conflict-free), and writes the R entries and
// their remainders to the stage's TMEM slot with tcgen05.st.  The running sum of what it read is the bias gradient -- the
// separate column-sum warp of ws_dw_kernel and its second read of the tile go away too.
//
//   warp 16      TMA producer (raw gzout / zin blocks of R rows x 32 columns, 32-byte-atom 128B swizzle)
//   warps 0-15   transform groups of 4 warps (one per TMEM lane quarter); group g takes ring iterations g, g + groups, ...:
//                gzout column -> TMEM (hi, lo) + column sum; zin tile: activation / dropout / branch scaling in place +
//                remainder tile; afterwards all 16 warps drain the accumulators to partial[split]
//   warp 17      MMA issuer: D += Ahi[tmem]*Bhi + Alo[tmem]*Bhi + Ahi[tmem]*Blo
//
// TMEM: mt*ntl*NT accumulator columns + stages * mt * 2R operand columns <= 512.
#pragma once

namespace pcfd {
namespace ws {

constexpr int DW1_W_TMA = 16, DW1_W_MMA = 17;
constexpr int DW1_THREADS = 18 * 32;

// byte offset of (row e, column c) inside a 32-column block of the 32-byte-atom 128B swizzle
__device__ __forceinline__ uint32_t atom32_off(uint32_t e, uint32_t c) {
  return e * 128u + ((((c >> 3) ^ (e & 3u)) << 5) | (((c >> 2) & 1u) << 4)) + ((c & 3u) << 2);
}

// activation / dropout / branch scaling of the staged zin tile in place + remainder tile (128 threads of one group)
template <int R, int ACT, bool DROP>
__device__ __forceinline__ void dw1_transform_b(uint8_t* bt, uint32_t lo_off, int b_items, int tt, const DwArgs& a,
                                                uint32_t hseed, int64_t row0, int k0) {
  constexpr int BLK = R * 128;
  const bool identity = ACT == PCFD_ACT_NONE && !DROP && a.tin.escale == nullptr;
#pragma unroll 2
  for (int idx = tt; idx < b_items; idx += 128) {
    const int j = idx & 7, r = (idx >> 3) & (R - 1), blk = idx / (8 * R);
    uint8_t* p = bt + blk * BLK + r * 128 + ((((uint32_t)(j >> 1) ^ (uint32_t)(r & 3)) << 5) | ((uint32_t)(j & 1) << 4));
    const float4 x = *reinterpret_cast<const float4*>(p);
    float v[1][4] = {{x.x, x.y, x.z, x.w}};
    const int col0 = k0 + blk * 32 + j * 4;
    if (!identity && col0 < a.tin.act_cols) {
      int64_t row = row0 + r;
      if (row >= a.rows) row = a.rows - 1;            // rows past the end are zero-filled by TMA; keep the indices valid
      const int64_t geom = a.tin.escale != nullptr ? geom_of(row, a.rows_per_geom) : 0;
      if (col0 + 4 <= a.tin.act_cols) {
        const uint32_t hrow = DROP ? dropout_row_hash(hseed, row) : 0u;
        const float* es = a.tin.escale != nullptr ? a.tin.escale + geom * a.tin.ldescale + col0 : nullptr;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float sc = 1.0f;
          if (DROP) sc = dropout_from_row(hrow, row, col0 + e, a.tin.drop_p, a.tin.inv_keep);
          if (es != nullptr) sc *= __ldg(es + e);
          float zz[1] = {v[0][e]};
          jet_act_fwd_t<1, ACT>(sc, zz);
          v[0][e] = zz[0];
        }
      } else {
        transform_dispatch<1>(v, a.tin, DROP || a.tin.escale != nullptr, hseed, row, geom, col0, a.tin.act_cols - col0);
      }
      if (row0 + r < a.rows) *reinterpret_cast<float4*>(p) = make_float4(v[0][0], v[0][1], v[0][2], v[0][3]);
      else { v[0][0] = v[0][1] = v[0][2] = v[0][3] = 0.0f; }
    }
    *reinterpret_cast<float4*>(p + lo_off) = make_float4(v[0][0] - trunc_tf32(v[0][0]), v[0][1] - trunc_tf32(v[0][1]),
                                                         v[0][2] - trunc_tf32(v[0][2]), v[0][3] - trunc_tf32(v[0][3]));
  }
}

template <int R, int NT>
__global__ void __launch_bounds__(DW1_THREADS, 1) ws_dw1_kernel(const __grid_constant__ CUtensorMap tmG,
                                                                const __grid_constant__ CUtensorMap tmZ, DwArgs a) {
  constexpr int BLK = R * 128;                    // one 32-column block of a stage
  constexpr int NB = NT / 32;
  static_assert(R % 16 == 0 && (R & (R - 1)) == 0 && BLK % 1024 == 0, "whole tcgen05.st groups, swizzle phase kept");

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
  const uint32_t B_HI = (uint32_t)AB * BLK, B_LO = B_HI + (uint32_t)BB * BLK;
  const uint32_t STAGE_BYTES = (uint32_t)(AB + 2 * BB) * BLK;
  const uint32_t A_TMEM = (uint32_t)(a.mt * a.ntl * NT);           // first operand column
  const int64_t r_begin = (int64_t)blockIdx.y * a.rows_per_split;
  const int64_t r_end = min(a.rows, r_begin + a.rows_per_split);
  const int nsteps = (int)((r_end - r_begin + R - 1) / R);

  if (tid == 0) {
    for (int s = 0; s < STG; ++s) {
      tc::mbar_init(&raw_full[s], 1);
      tc::mbar_init(&ops_ready[s], 128);
      tc::mbar_init(&stage_free[s], 1);
    }
    tc::mbar_init(&acc_full, 1);
    tc::fence_mbar_init();
  }
  if (warp == DW1_W_MMA) tc::tmem_alloc(&tmem_base_s, 512);
  if (warp == DW1_W_TMA && lane == 0) { prefetch_tmap(&tmG); prefetch_tmap(&tmZ); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  griddep_wait();                  // everything above overlaps the tail of the previous kernel of the stream
  griddep_launch_dependents();     // one resident wave: the next kernel may take SMs as they free up

  if (warp == DW1_W_TMA) {
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
        for (int b = 0; b < ab; ++b) tma_load_2d(st + b * BLK, &tmG, n0 + b * 32, row, &raw_full[s]);
        for (int b = 0; b < bb; ++b) tma_load_2d(st + B_HI + b * BLK, &tmZ, k0 + b * 32, row, &raw_full[s]);
      }
      __syncwarp();
      if (++s == STG) { s = 0; ph ^= 1; }
    }
  } else if (warp == DW1_W_MMA) {
    // ================================ MMA issuer ================================
    constexpr uint32_t IDESC = tc::make_idesc_tf32(128, NT, false, true);      // A from TMEM, B MN-major
    const uint64_t dbase = desc_mnmajor(tc::smem_u32(smem), BLK, 512);        // + (byte offset >> 4) in the address field
    const uint32_t b_hi = B_HI >> 4, b_lo = B_LO >> 4;
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nsteps; ++it) {
      tc::bounded_wait(&ops_ready[s], ph);
      tc::tc_fence_after();
      const uint64_t ds = dbase + (uint64_t)(((uint32_t)s * STAGE_BYTES) >> 4);
      const uint32_t ta = tmem_base + A_TMEM + (uint32_t)(s * a.mt) * (2 * R);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < R / 8; ++ks) {
          for (int mt = 0; mt < a.mt; ++mt) {
            const uint32_t a_hi = ta + (uint32_t)mt * (2 * R) + ks * 8, a_lo = a_hi + R;
            for (int nl = 0; nl < a.ntl; ++nl) {
              const uint64_t db_hi = ds + b_hi + (uint32_t)((nl * NB * BLK + ks * 1024) >> 4);
              const uint64_t db_lo = db_hi + (b_lo - b_hi);
              const uint32_t d = tmem_base + (uint32_t)(mt * a.ntl + nl) * NT;
              mma_tf32_ts(d, a_hi, db_hi, IDESC, (it > 0 || ks > 0) ? 1u : 0u);
              mma_tf32_ts(d, a_lo, db_hi, IDESC, 1u);
              mma_tf32_ts(d, a_hi, db_lo, IDESC, 1u);
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
  } else {
    // ================================ transform ================================
    const int g = warp >> 2, q = warp & 3;            // group, TMEM lane quarter
    const int tt = q * 32 + lane;                     // 0..127 within the group = TMEM lane = gzout column of an m-tile
    const uint64_t seed = a.tin.seed_dev ? *a.tin.seed_dev : 0ULL;
    const uint32_t hseed = dropout_seed_hash(seed, a.tin.salt);
    const bool drop = a.tin.drop_p > 0.0f;
    const int b_items = bb * R * 8;                   // (block, row, 16-byte chunk) positions of the zin tile
    float csum[4] = {0.f, 0.f, 0.f, 0.f};             // column sums of this thread's gzout columns (one per m-tile)
    int s = g;                                        // g < groups <= stages
    uint32_t ph = 0;
    for (int it = g; it < (g < a.groups ? nsteps : 0); it += a.groups) {
      uint8_t* st = smem + (size_t)s * STAGE_BYTES;
      const int64_t row0 = r_begin + (int64_t)it * R;
      tc::bounded_wait(&raw_full[s], ph);
      tc::tc_fence_after();     // the MMAs that read this stage's TMEM slot completed before its refill was issued
      // ---- A: column tt of every m-tile -> TMEM (hi = the fp32 word, lo = exact remainder), running column sum
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        if (mt < a.mt && mt * 4 + q < ab) {
          const uint8_t* blk = st + (mt * 4 + q) * BLK;
          const uint32_t tcol = tmem_base + ((uint32_t)(32 * q) << 16) + A_TMEM + (uint32_t)(s * a.mt + mt) * (2 * R);
#pragma unroll
          for (int e0 = 0; e0 < R; e0 += 16) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float x = *reinterpret_cast<const float*>(blk + atom32_off((uint32_t)(e0 + e), (uint32_t)lane));
              csum[mt] += x;
              hi[e] = __float_as_uint(x);
              lo[e] = __float_as_uint(x - trunc_tf32(x));
            }
            tmem_st16(tcol + e0, hi);
            tmem_st16(tcol + R + e0, lo);
          }
        }
      }
      // ---- B: activation in place + remainder tile
      uint8_t* bt = st + B_HI;
      const uint32_t lo_off = B_LO - B_HI;
      if (a.tin.act == PCFD_ACT_SILU) {
        if (drop) dw1_transform_b<R, PCFD_ACT_SILU, true>(bt, lo_off, b_items, tt, a, hseed, row0, k0);
        else dw1_transform_b<R, PCFD_ACT_SILU, false>(bt, lo_off, b_items, tt, a, hseed, row0, k0);
      } else if (a.tin.act == PCFD_ACT_TANH) {
        if (drop) dw1_transform_b<R, PCFD_ACT_TANH, true>(bt, lo_off, b_items, tt, a, hseed, row0, k0);
        else dw1_transform_b<R, PCFD_ACT_TANH, false>(bt, lo_off, b_items, tt, a, hseed, row0, k0);
      } else {
        if (drop) dw1_transform_b<R, PCFD_ACT_NONE, true>(bt, lo_off, b_items, tt, a, hseed, row0, k0);
        else dw1_transform_b<R, PCFD_ACT_NONE, false>(bt, lo_off, b_items, tt, a, hseed, row0, k0);
      }
      tmem_st_wait();
      tc::fence_proxy_async();
      tc::tc_fence_before();
      mbar_arrive(&ops_ready[s]);
      s += a.groups;
      if (s >= STG) { s -= STG; ph ^= 1; }
    }

    // ================================ epilogue ================================
    if (nsteps > 0) {
      tc::bounded_wait(&acc_full, 0);        // all MMAs done: the ring is free to carry the column sums
      tc::tc_fence_after();
    }
    // ---- bias gradient: the groups' column sums, added in group order
    if (a.colsum != nullptr && pk == 0) {
      float* cs = reinterpret_cast<float*>(smem);       // [group][m-tile][128]
      if (g < a.groups) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
          if (mt < a.mt) cs[(g * 4 + mt) * 128 + tt] = csum[mt];
      }
      named_bar_sync(1, 512);
      if (g == 0) {
        float* dst = a.colsum + (int64_t)blockIdx.y * a.n;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          const int col = n0 + mt * 128 + tt;
          if (mt < a.mt && col < a.n) {
            float v = 0.0f;
            for (int gg = 0; gg < a.groups; ++gg) v += cs[(gg * 4 + mt) * 128 + tt];
            dst[col] = v;
          }
        }
      }
    }
    // ---- accumulators -> partial[split]
    const int cgrp = warp >> 2;                       // column blocks cgrp, cgrp + 4, ...
    const int blocks_per_mt = a.ntl * NB;
    const int total_cb = a.mt * blocks_per_mt;
    float* dst = a.partial + (int64_t)blockIdx.y * a.n * a.k;
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
  if (warp == DW1_W_MMA) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace ws
}  // namespace pcfd
