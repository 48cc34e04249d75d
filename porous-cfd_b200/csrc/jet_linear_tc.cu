// Jet linear layer on the 5th-generation tensor cores (engine 1): tcgen05.mma kind::tf32 with
// 3xTF32 split operands, accumulators in TMEM.
//
// One CTA = 128 points x all cj channels x NT output columns.  Each (channel) owns a 128 x NT fp32
// accumulator in TMEM (cj*NT <= 512 columns).  Per contraction step of BK inputs the 256 threads
//   1. load the pre-activation jets of their row from HBM (float4),
//   2. apply the input transform (activation jet, dropout mask, branch scaling) in registers,
//   3. split into TF32 hi/lo and store 16-byte chunks into the UMMA no-swizzle K-major layout,
//   4. fence.proxy.async + barrier; one thread issues 3 x cj x BK/8 tcgen05.mma and commits them
//      to the stage's mbarrier, which frees the stage two steps later (double buffering: the
//      transform of step i+1 overlaps the MMAs of step i).
// Epilogue: tcgen05.ld (32 lanes x 16 columns per warp), bias / per-geometry constant on the value
// channel, float4 stores of the new pre-activations.
#include "common.cuh"
#include "tc_common.cuh"

namespace pcfd {

struct TcFwdArgs {
  const float* zin; int64_t zin_ps; int ldzin;
  InTrans tin;
  const float* w; int ldw; const float* bias; const float* cvec; int ldcvec;
  float* zout; int64_t zout_ps; int ldzout;
  int64_t rows, rows_per_geom; int k, n;
  int vec_in, vec_w, vec_out;
};

using tc::bounded_wait;

template <int CJ, int BK, int NT>
__global__ void __launch_bounds__(256, 1) jet_fwd_tc_kernel(TcFwdArgs a) {
  constexpr int STAGES = 2;
  constexpr int KCH = BK / 4;                 // 16-byte chunks per row and stage
  constexpr int CPT = KCH / 2;                // chunks per thread (two threads share a row)
  constexpr int A_TILE = 128 * BK * 4;
  constexpr int B_TILE = NT * BK * 4;
  constexpr int STAGE_BYTES = CJ * 2 * A_TILE + 2 * B_TILE;
  constexpr uint32_t LBO_A = 128 * 16, LBO_B = NT * 16, SBO = 128;
  constexpr int B_CHUNKS = NT * KCH;          // 16-byte chunks of one B tile
  constexpr int BPT = (B_CHUNKS + 255) / 256;
  constexpr uint32_t NEED = CJ * NT;
  constexpr uint32_t TMEM_COLS = NEED <= 32 ? 32 : NEED <= 64 ? 64 : NEED <= 128 ? 128 : NEED <= 256 ? 256 : 512;
  static_assert(NEED <= 512, "accumulators exceed tensor memory");
  static_assert(CPT >= 1, "BK must be at least 8");

  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mma_done[STAGES];
  __shared__ __align__(8) uint64_t acc_done;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row_l = tid & 127, half = tid >> 7;
  const int64_t row0 = (int64_t)blockIdx.x * 128;
  const int n0 = blockIdx.y * NT;
  const int64_t row = row0 + row_l;
  const bool valid = row < a.rows;
  const int64_t geom = valid ? geom_of(row, a.rows_per_geom) : 0;
  const uint64_t seed = a.tin.seed_dev ? *a.tin.seed_dev : 0ULL;
  const bool plain = (a.tin.act == PCFD_ACT_NONE && a.tin.escale == nullptr && a.tin.drop_p == 0.0f);

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) tc::mbar_init(&mma_done[s], 1);
    tc::mbar_init(&acc_done, 1);
    tc::fence_mbar_init();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  constexpr uint32_t IDESC = tc::make_idesc_tf32(128, NT, false, false);

  const int nchunks = (a.k + BK - 1) / BK;
  // two register buffers: the HBM loads of chunk i+2 are issued while chunk i is transformed (two chunks,
  // ~64 KB per SM, stay in flight -- enough to cover the memory latency at full HBM bandwidth)
  float zA[CPT][CJ][4], zB[CPT][CJ][4];
  float4 wA[BPT], wB[BPT];

  auto load_chunk = [&](int i, float (&z)[CPT][CJ][4], float4 (&wv)[BPT]) {
    const int k0 = i * BK;
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
      const int kb = k0 + (half * CPT + q) * 4;
#pragma unroll
      for (int c = 0; c < CJ; ++c) {
        const float* src = a.zin + c * a.zin_ps + row * a.ldzin + kb;
        if (valid && a.vec_in && kb + 3 < a.k) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src));
          z[q][c][0] = v.x; z[q][c][1] = v.y; z[q][c][2] = v.z; z[q][c][3] = v.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) z[q][c][e] = (valid && kb + e < a.k) ? __ldg(src + e) : 0.0f;
        }
      }
    }
#pragma unroll
    for (int t = 0; t < BPT; ++t) {
      const int idx = tid + t * 256;
      wv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < B_CHUNKS) {
        const int nr = n0 + idx % NT, kb = k0 + (idx / NT) * 4;
        if (nr < a.n) {
          const float* src = a.w + (int64_t)nr * a.ldw + kb;
          if (a.vec_w && kb + 3 < a.k) {
            wv[t] = __ldg(reinterpret_cast<const float4*>(src));
          } else {
            wv[t].x = kb + 0 < a.k ? __ldg(src + 0) : 0.f;
            wv[t].y = kb + 1 < a.k ? __ldg(src + 1) : 0.f;
            wv[t].z = kb + 2 < a.k ? __ldg(src + 2) : 0.f;
            wv[t].w = kb + 3 < a.k ? __ldg(src + 3) : 0.f;
          }
        }
      }
    }
  };

  auto stage_chunk = [&](int i, uint8_t* st, float (&z)[CPT][CJ][4], float4 (&wv)[BPT]) {
    const int k0 = i * BK;
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
      const int j = half * CPT + q;
      const int kb = k0 + j * 4;
      if (!plain && valid) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = kb + e;
          if (col < a.tin.act_cols && col < a.k) {
            float zz[CJ];
#pragma unroll
            for (int c = 0; c < CJ; ++c) zz[c] = z[q][c][e];
            float m;
            const float s = in_scale(a.tin, seed, row, geom, col, m);
            jet_act_fwd<CJ>(a.tin.act, s, zz);
#pragma unroll
            for (int c = 0; c < CJ; ++c) z[q][c][e] = zz[c];
          }
        }
      }
      const uint32_t off = j * LBO_A + (row_l >> 3) * SBO + (row_l & 7) * 16;
#pragma unroll
      for (int c = 0; c < CJ; ++c) {
        float4 hi, lo;
        tc::split4(z[q][c], hi, lo);
        *reinterpret_cast<float4*>(st + (2 * c) * A_TILE + off) = hi;
        *reinterpret_cast<float4*>(st + (2 * c + 1) * A_TILE + off) = lo;
      }
    }
    uint8_t* bt = st + CJ * 2 * A_TILE;
#pragma unroll
    for (int t = 0; t < BPT; ++t) {
      const int idx = tid + t * 256;
      if (idx < B_CHUNKS) {
        const int nr = idx % NT, j = idx / NT;
        const float v[4] = {wv[t].x, wv[t].y, wv[t].z, wv[t].w};
        float4 hi, lo;
        tc::split4(v, hi, lo);
        const uint32_t off = j * LBO_B + (nr >> 3) * SBO + (nr & 7) * 16;
        *reinterpret_cast<float4*>(bt + off) = hi;
        *reinterpret_cast<float4*>(bt + B_TILE + off) = lo;
      }
    }
  };

  auto process = [&](int i, float (&z)[CPT][CJ][4], float4 (&wv)[BPT]) {
    const int s = i & 1;
    uint8_t* st = smem + s * STAGE_BYTES;
    if (i >= 2) bounded_wait(&mma_done[s], ((i >> 1) - 1) & 1);     // MMAs that read this stage are done
    stage_chunk(i, st, z, wv);
    if (i + 2 < nchunks) load_chunk(i + 2, z, wv);
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      const uint32_t sbase = tc::smem_u32(st);
      const uint32_t bbase = sbase + CJ * 2 * A_TILE;
#pragma unroll
      for (int ks = 0; ks < BK / 8; ++ks) {
        const uint64_t db_hi = tc::make_smem_desc(bbase + ks * 2 * LBO_B, LBO_B, SBO);
        const uint64_t db_lo = tc::make_smem_desc(bbase + B_TILE + ks * 2 * LBO_B, LBO_B, SBO);
#pragma unroll
        for (int c = 0; c < CJ; ++c) {
          const uint64_t da_hi = tc::make_smem_desc(sbase + (2 * c) * A_TILE + ks * 2 * LBO_A, LBO_A, SBO);
          const uint64_t da_lo = tc::make_smem_desc(sbase + (2 * c + 1) * A_TILE + ks * 2 * LBO_A, LBO_A, SBO);
          const uint32_t d = tmem_base + c * NT;
          tc::mma_tf32(d, da_hi, db_hi, IDESC, (i > 0 || ks > 0) ? 1u : 0u);
          tc::mma_tf32(d, da_lo, db_hi, IDESC, 1u);
          tc::mma_tf32(d, da_hi, db_lo, IDESC, 1u);
        }
      }
      tc::mma_commit(&mma_done[s]);
      if (i + 1 == nchunks) tc::mma_commit(&acc_done);
    }
  };
  load_chunk(0, zA, wA);
  if (nchunks > 1) load_chunk(1, zB, wB);
  for (int i = 0; i < nchunks; i += 2) {
    process(i, zA, wA);
    if (i + 1 < nchunks) process(i + 1, zB, wB);
  }

  bounded_wait(&acc_done, 0);
  tc::tc_fence_after();

  // epilogue: warp (q = warp & 3) owns TMEM lanes [32q, 32q+32); the two warpgroups split the columns
  const int q = warp & 3, hcol = warp >> 2;
  const int64_t orow = row0 + 32 * q + lane;
  const bool ovalid = orow < a.rows;
  const int64_t ogeom = ovalid ? geom_of(orow, a.rows_per_geom) : 0;
#pragma unroll
  for (int c = 0; c < CJ; ++c) {
#pragma unroll
    for (int cb = 0; cb < NT / 32; ++cb) {
      const int col = hcol * (NT / 2) + cb * 16;
      float v[16];
      tc::tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + c * NT + col, v);
      const int nc = n0 + col;
      if (ovalid && nc < a.n) {
        if (c == 0) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            if (nc + e < a.n) {
              if (a.bias != nullptr) v[e] += __ldg(a.bias + nc + e);
              if (a.cvec != nullptr) v[e] += __ldg(a.cvec + ogeom * a.ldcvec + nc + e);
            }
          }
        }
        float* dst = a.zout + c * a.zout_ps + orow * a.ldzout + nc;
        if (a.vec_out && nc + 15 < a.n) {
#pragma unroll
          for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(dst + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) if (nc + e < a.n) dst[e] = v[e];
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int CJ, int BK, int NT>
static int launch_fwd_tc(const TcFwdArgs& a, cudaStream_t st) {
  constexpr int SMEM = 2 * (CJ * 2 * 128 * BK * 4 + 2 * NT * BK * 4);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(jet_fwd_tc_kernel<CJ, BK, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;
    configured = true;
  }
  dim3 grid((unsigned)((a.rows + 127) / 128), (unsigned)((a.n + NT - 1) / NT));
  jet_fwd_tc_kernel<CJ, BK, NT><<<grid, 256, SMEM, st>>>(a);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace pcfd

using namespace pcfd;

extern "C" int pcfd_tc_supported_fwd(int32_t cj, int64_t rows, int32_t k, int32_t n, int32_t ldzin, int32_t ldw,
                                     int32_t ldzout) {
  (void)ldzin; (void)ldw; (void)ldzout;
  return valid_cj(cj) && rows >= 512 && (int64_t)k * n >= 64;
}

extern "C" int pcfd_tc_jet_linear_fwd(const float* zin, int64_t zin_ps, int32_t ldzin, const pcfd_intrans_t* tin,
                                      const float* w, int32_t ldw, const float* bias, const float* cvec,
                                      int32_t ldcvec, float* zout, int64_t zout_ps, int32_t ldzout, int32_t cj,
                                      int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n, void* stream) {
  TcFwdArgs a{zin, zin_ps, ldzin, make_intrans(tin, k), w, ldw, bias, cvec, ldcvec, zout, zout_ps, ldzout,
              rows, rows_per_geom, k, n, 0, 0, 0};
  a.vec_in = al16(zin) && ldzin % 4 == 0 && zin_ps % 4 == 0;
  a.vec_w = al16(w) && ldw % 4 == 0;
  a.vec_out = al16(zout) && ldzout % 4 == 0 && zout_ps % 4 == 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (cj) {
    case 1: return n > 128 ? launch_fwd_tc<1, 16, 256>(a, st) : (n > 32 ? launch_fwd_tc<1, 16, 128>(a, st) : launch_fwd_tc<1, 16, 32>(a, st));
    case 3: return n > 32 ? launch_fwd_tc<3, 16, 128>(a, st) : launch_fwd_tc<3, 16, 32>(a, st);
    case 4: return n > 32 ? launch_fwd_tc<4, 16, 128>(a, st) : launch_fwd_tc<4, 16, 32>(a, st);
    case 5: return n > 32 ? launch_fwd_tc<5, 16, 64>(a, st) : launch_fwd_tc<5, 16, 32>(a, st);
    case 7: return n > 32 ? launch_fwd_tc<7, 8, 64>(a, st) : launch_fwd_tc<7, 8, 32>(a, st);
  }
  return PCFD_ERR_ARG;
}

// =================================================================================================
// backward to the layer input:  gzin[c][row][kout] = reverse_transform( sum_n gzout[c][row][n] * W[n][kout] )
// A = gzout tile (K-major, contraction over the layer's outputs).  B[kout][n] must also be K-major
// (tcgen05 kind::tf32 returned zeros for MN-major no-swizzle operands on this part -- see
// scripts/probe/mn_probe.cu), so the W tile is transposed on the way in: a thread gathers 4
// consecutive contraction rows of one kout column (loads coalesced across kout) into one chunk.
// =================================================================================================
namespace pcfd {

struct TcDxArgs {
  const float* gzout; int64_t gzout_ps; int ldgzout;
  const float* w; int ldw;
  const float* zin; int64_t zin_ps; int ldzin;
  InTrans tin;
  float* gzin; int64_t gzin_ps; int ldgzin;
  float* gescale; int ldgescale;
  int64_t rows, rows_per_geom; int k, n;
  int vec_g, vec_w, vec_z, vec_out;
};

template <int EC> struct TmemLd;
template <> struct TmemLd<16> { static __device__ __forceinline__ void ld(uint32_t a, float (&v)[16]) { tc::tmem_ld16(a, v); } };
template <> struct TmemLd<8> { static __device__ __forceinline__ void ld(uint32_t a, float (&v)[8]) { tc::tmem_ld8(a, v); } };

template <int CJ, int BK, int NT, int EC>
__global__ void __launch_bounds__(256, 1) jet_dx_tc_kernel(TcDxArgs a) {
  constexpr int STAGES = 2;
  constexpr int KCH = BK / 4, CPT = KCH / 2;
  constexpr int A_TILE = 128 * BK * 4;
  constexpr int B_TILE = NT * BK * 4;
  constexpr int STAGE_BYTES = CJ * 2 * A_TILE + 2 * B_TILE;
  constexpr uint32_t LBO_A = 128 * 16, SBO = 128;
  constexpr uint32_t LBO_B = NT * 16;
  constexpr int B_CHUNKS = NT * KCH;
  constexpr int BPT = (B_CHUNKS + 255) / 256;
  constexpr uint32_t NEED = CJ * NT;
  constexpr uint32_t TMEM_COLS = NEED <= 32 ? 32 : NEED <= 64 ? 64 : NEED <= 128 ? 128 : NEED <= 256 ? 256 : 512;
  static_assert(NEED <= 512, "accumulators exceed tensor memory");

  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mma_done[STAGES];
  __shared__ __align__(8) uint64_t acc_done;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row_l = tid & 127, half = tid >> 7;
  const int64_t row0 = (int64_t)blockIdx.x * 128;
  const int c0 = blockIdx.y * NT;                  // first input column (kout) of this tile
  const int64_t row = row0 + row_l;
  const bool valid = row < a.rows;
  const uint64_t seed = a.tin.seed_dev ? *a.tin.seed_dev : 0ULL;
  const bool plain = (a.tin.act == PCFD_ACT_NONE && a.tin.escale == nullptr && a.tin.drop_p == 0.0f);

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) tc::mbar_init(&mma_done[s], 1);
    tc::mbar_init(&acc_done, 1);
    tc::fence_mbar_init();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  constexpr uint32_t IDESC = tc::make_idesc_tf32(128, NT, false, false);

  const int nchunks = (a.n + BK - 1) / BK;
  float gA[CPT][CJ][4], gB[CPT][CJ][4];          // two register buffers: loads run two chunks ahead
  float4 wA[BPT], wB[BPT];

  auto load_chunk = [&](int i, float (&g)[CPT][CJ][4], float4 (&wv)[BPT]) {
    const int nb0 = i * BK;
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
      const int nb = nb0 + (half * CPT + q) * 4;
#pragma unroll
      for (int c = 0; c < CJ; ++c) {
        const float* src = a.gzout + c * a.gzout_ps + row * a.ldgzout + nb;
        if (valid && a.vec_g && nb + 3 < a.n) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src));
          g[q][c][0] = v.x; g[q][c][1] = v.y; g[q][c][2] = v.z; g[q][c][3] = v.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) g[q][c][e] = (valid && nb + e < a.n) ? __ldg(src + e) : 0.0f;
        }
      }
    }
#pragma unroll
    for (int t = 0; t < BPT; ++t) {
      const int idx = tid + t * 256;
      wv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < B_CHUNKS) {
        const int col = c0 + idx % NT, nr = nb0 + (idx / NT) * 4;     // chunk = W[nr..nr+3][col]
        if (col < a.k) {
          const float* src = a.w + (int64_t)nr * a.ldw + col;
          wv[t].x = nr + 0 < a.n ? __ldg(src) : 0.f;
          wv[t].y = nr + 1 < a.n ? __ldg(src + a.ldw) : 0.f;
          wv[t].z = nr + 2 < a.n ? __ldg(src + 2 * (int64_t)a.ldw) : 0.f;
          wv[t].w = nr + 3 < a.n ? __ldg(src + 3 * (int64_t)a.ldw) : 0.f;
        }
      }
    }
  };
  auto stage_chunk = [&](uint8_t* st, float (&g)[CPT][CJ][4], float4 (&wv)[BPT]) {
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
      const int j = half * CPT + q;
      const uint32_t off = j * LBO_A + (row_l >> 3) * SBO + (row_l & 7) * 16;
#pragma unroll
      for (int c = 0; c < CJ; ++c) {
        float4 hi, lo;
        tc::split4(g[q][c], hi, lo);
        *reinterpret_cast<float4*>(st + (2 * c) * A_TILE + off) = hi;
        *reinterpret_cast<float4*>(st + (2 * c + 1) * A_TILE + off) = lo;
      }
    }
    uint8_t* bt = st + CJ * 2 * A_TILE;
#pragma unroll
    for (int t = 0; t < BPT; ++t) {
      const int idx = tid + t * 256;
      if (idx < B_CHUNKS) {
        const int cl = idx % NT, j = idx / NT;
        const float v[4] = {wv[t].x, wv[t].y, wv[t].z, wv[t].w};
        float4 hi, lo;
        tc::split4(v, hi, lo);
        const uint32_t off = j * LBO_B + (cl >> 3) * SBO + (cl & 7) * 16;
        *reinterpret_cast<float4*>(bt + off) = hi;
        *reinterpret_cast<float4*>(bt + B_TILE + off) = lo;
      }
    }
  };

  auto process = [&](int i, float (&g)[CPT][CJ][4], float4 (&wv)[BPT]) {
    const int s = i & 1;
    uint8_t* st = smem + s * STAGE_BYTES;
    if (i >= 2) bounded_wait(&mma_done[s], ((i >> 1) - 1) & 1);
    stage_chunk(st, g, wv);
    if (i + 2 < nchunks) load_chunk(i + 2, g, wv);
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      const uint32_t sbase = tc::smem_u32(st);
      const uint32_t bbase = sbase + CJ * 2 * A_TILE;
#pragma unroll
      for (int ks = 0; ks < BK / 8; ++ks) {
        const uint64_t db_hi = tc::make_smem_desc(bbase + ks * 2 * LBO_B, LBO_B, SBO);
        const uint64_t db_lo = tc::make_smem_desc(bbase + B_TILE + ks * 2 * LBO_B, LBO_B, SBO);
#pragma unroll
        for (int c = 0; c < CJ; ++c) {
          const uint64_t da_hi = tc::make_smem_desc(sbase + (2 * c) * A_TILE + ks * 2 * LBO_A, LBO_A, SBO);
          const uint64_t da_lo = tc::make_smem_desc(sbase + (2 * c + 1) * A_TILE + ks * 2 * LBO_A, LBO_A, SBO);
          const uint32_t d = tmem_base + c * NT;
          tc::mma_tf32(d, da_hi, db_hi, IDESC, (i > 0 || ks > 0) ? 1u : 0u);
          tc::mma_tf32(d, da_lo, db_hi, IDESC, 1u);
          tc::mma_tf32(d, da_hi, db_lo, IDESC, 1u);
        }
      }
      tc::mma_commit(&mma_done[s]);
      if (i + 1 == nchunks) tc::mma_commit(&acc_done);
    }
  };
  load_chunk(0, gA, wA);
  if (nchunks > 1) load_chunk(1, gB, wB);
  for (int i = 0; i < nchunks; i += 2) {
    process(i, gA, wA);
    if (i + 1 < nchunks) process(i + 1, gB, wB);
  }

  bounded_wait(&acc_done, 0);
  tc::tc_fence_after();

  // epilogue: reverse input transform on whole jets, store d/dzin, accumulate d/descale
  const int q = warp & 3, hcol = warp >> 2;
  const int64_t orow = row0 + 32 * q + lane;
  const bool ovalid = orow < a.rows;
  const int64_t ogeom = ovalid ? geom_of(orow, a.rows_per_geom) : 0;
  const int64_t geom0 = __shfl_sync(0xffffffffu, ogeom, 0);
  const bool uniform = __all_sync(0xffffffffu, (!ovalid) || ogeom == geom0);
#pragma unroll 1
  for (int cb = 0; cb < (NT / 2) / EC; ++cb) {
    const int col = hcol * (NT / 2) + cb * EC;
    const int kc = c0 + col;
    float acc[CJ][EC];
#pragma unroll
    for (int c = 0; c < CJ; ++c) TmemLd<EC>::ld(tmem_base + ((uint32_t)(32 * q) << 16) + c * NT + col, acc[c]);
    float ge[EC];
#pragma unroll
    for (int e = 0; e < EC; ++e) ge[e] = 0.0f;
    if (ovalid && kc < a.k) {
      if (!plain) {
        float z[CJ][EC];
#pragma unroll
        for (int c = 0; c < CJ; ++c) {
          const float* src = a.zin + c * a.zin_ps + orow * a.ldzin + kc;
          if (a.vec_z && kc + EC - 1 < a.k) {
#pragma unroll
            for (int e = 0; e < EC; e += 4) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(src + e));
              z[c][e] = v.x; z[c][e + 1] = v.y; z[c][e + 2] = v.z; z[c][e + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int e = 0; e < EC; ++e) z[c][e] = kc + e < a.k ? __ldg(src + e) : 0.0f;
          }
        }
#pragma unroll
        for (int e = 0; e < EC; ++e) {
          const int cc = kc + e;
          if (cc < a.k && cc < a.tin.act_cols) {
            float gg[CJ], zz[CJ];
#pragma unroll
            for (int c = 0; c < CJ; ++c) { gg[c] = acc[c][e]; zz[c] = z[c][e]; }
            float m;
            const float s = in_scale(a.tin, seed, orow, ogeom, cc, m);
            ge[e] = jet_act_bwd<CJ>(a.tin.act, s, m, zz, gg);
#pragma unroll
            for (int c = 0; c < CJ; ++c) acc[c][e] = gg[c];
          }
        }
      }
#pragma unroll
      for (int c = 0; c < CJ; ++c) {
        float* dst = a.gzin + c * a.gzin_ps + orow * a.ldgzin + kc;
        if (a.vec_out && kc + EC - 1 < a.k) {
#pragma unroll
          for (int e = 0; e < EC; e += 4)
            *reinterpret_cast<float4*>(dst + e) = make_float4(acc[c][e], acc[c][e + 1], acc[c][e + 2], acc[c][e + 3]);
        } else {
#pragma unroll
          for (int e = 0; e < EC; ++e) if (kc + e < a.k) dst[e] = acc[c][e];
        }
      }
    }
    if (a.gescale != nullptr) {
      if (uniform) {
#pragma unroll
        for (int e = 0; e < EC; ++e) {
          float v = ge[e];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0 && kc + e < a.k && v != 0.0f) atomicAdd(a.gescale + geom0 * a.ldgescale + kc + e, v);
        }
      } else if (ovalid) {
#pragma unroll
        for (int e = 0; e < EC; ++e)
          if (kc + e < a.k && ge[e] != 0.0f) atomicAdd(a.gescale + ogeom * a.ldgescale + kc + e, ge[e]);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int CJ, int BK, int NT, int EC>
static int launch_dx_tc(const TcDxArgs& a, cudaStream_t st) {
  constexpr int SMEM = 2 * (CJ * 2 * 128 * BK * 4 + 2 * NT * BK * 4);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(jet_dx_tc_kernel<CJ, BK, NT, EC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;
    configured = true;
  }
  dim3 grid((unsigned)((a.rows + 127) / 128), (unsigned)((a.k + NT - 1) / NT));
  jet_dx_tc_kernel<CJ, BK, NT, EC><<<grid, 256, SMEM, st>>>(a);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

// =================================================================================================
// backward to the weights:  partial[split][n][k] = sum over (channel, row in split) gzout[c][row][n] * T(zin)[c][row][k]
// =================================================================================================
struct TcDwArgs {
  const float* gzout; int64_t gzout_ps; int ldgzout;
  const float* zin; int64_t zin_ps; int ldzin;
  InTrans tin;
  float* partial;
  int64_t rows, rows_per_geom, rows_per_split; int k, n, tiles_k;
  int vec_g, vec_z;
};

// Contraction index of a stage: e = c*BR + p (channel-major over BR points).  Both operands are
// K-major tiles whose 16-byte chunks hold 4 consecutive contraction entries (4 consecutive points of
// one channel) of one output row; the loads that build a chunk are strided by the row pitch but
// coalesced across the threads of a warp (consecutive n / k columns).
template <int CJ, int BR, int NT>
__global__ void __launch_bounds__(256, 1) jet_dw_tc_kernel(TcDwArgs a) {
  constexpr int STAGES = 2;
  constexpr int E = CJ * BR;                       // contraction entries per stage
  static_assert(E % 8 == 0 && BR % 4 == 0, "stage contraction length must be a multiple of the MMA K");
  constexpr int A_TILE = E * 128 * 4;
  constexpr int B_TILE = E * NT * 4;
  constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;
  constexpr uint32_t SBO = 128, LBO_A = 128 * 16, LBO_B = NT * 16;
  constexpr int A_ITEMS = 128 * (E / 4);           // chunks of the gzout^T tile
  constexpr int APT = A_ITEMS / 256;
  constexpr int B_ITEMS = NT * (BR / 4);           // (k column, group of 4 points) items, all channels each
  constexpr int BPT = (B_ITEMS + 255) / 256;
  constexpr uint32_t TMEM_COLS = NT <= 32 ? 32 : NT <= 64 ? 64 : NT <= 128 ? 128 : 256;

  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mma_done[STAGES];
  __shared__ __align__(8) uint64_t acc_done;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile_n = blockIdx.x / a.tiles_k, tile_k = blockIdx.x % a.tiles_k;
  const int n0 = tile_n * 128, k0 = tile_k * NT;
  const int64_t r_begin = (int64_t)blockIdx.y * a.rows_per_split;
  const int64_t r_end = min(a.rows, r_begin + a.rows_per_split);
  const uint64_t seed = a.tin.seed_dev ? *a.tin.seed_dev : 0ULL;
  const bool plain = (a.tin.act == PCFD_ACT_NONE && a.tin.escale == nullptr && a.tin.drop_p == 0.0f);

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) tc::mbar_init(&mma_done[s], 1);
    tc::mbar_init(&acc_done, 1);
    tc::fence_mbar_init();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  constexpr uint32_t IDESC = tc::make_idesc_tf32(128, NT, false, false);

  float gvA[APT][4], gvB[APT][4];                  // two register buffers: loads run two stages ahead
  float zvA[BPT][CJ][4], zvB[BPT][CJ][4];          // [item][channel][point within the group of 4]

  auto load_chunk = [&](int64_t r0, float (&gv)[APT][4], float (&zv)[BPT][CJ][4]) {
#pragma unroll
    for (int t = 0; t < APT; ++t) {
      const int idx = tid + t * 256;
      const int nl = idx & 127, j = idx >> 7;          // chunk j = entries 4j..4j+3
      const int c = (4 * j) / BR, p = (4 * j) % BR;
      const int col = n0 + nl;
      const float* src = a.gzout + c * a.gzout_ps + (r0 + p) * a.ldgzout + col;
#pragma unroll
      for (int e = 0; e < 4; ++e) gv[t][e] = (col < a.n && r0 + p + e < r_end) ? __ldg(src + (int64_t)e * a.ldgzout) : 0.0f;
    }
#pragma unroll
    for (int t = 0; t < BPT; ++t) {
      const int idx = tid + t * 256;
      const int kl = idx % NT, pj = idx / NT;
      const int col = k0 + kl;
      const bool ok = idx < B_ITEMS && col < a.k;
#pragma unroll
      for (int c = 0; c < CJ; ++c) {
        const float* src = a.zin + c * a.zin_ps + (r0 + 4 * pj) * a.ldzin + col;
#pragma unroll
        for (int e = 0; e < 4; ++e) zv[t][c][e] = (ok && r0 + 4 * pj + e < r_end) ? __ldg(src + (int64_t)e * a.ldzin) : 0.0f;
      }
    }
  };
  auto stage_chunk = [&](int64_t r0, uint8_t* st, float (&gv)[APT][4], float (&zv)[BPT][CJ][4]) {
#pragma unroll
    for (int t = 0; t < APT; ++t) {
      const int idx = tid + t * 256;
      const int nl = idx & 127, j = idx >> 7;
      float4 hi, lo;
      tc::split4(gv[t], hi, lo);
      const uint32_t off = j * LBO_A + (nl >> 3) * SBO + (nl & 7) * 16;
      *reinterpret_cast<float4*>(st + off) = hi;
      *reinterpret_cast<float4*>(st + A_TILE + off) = lo;
    }
    uint8_t* bt = st + 2 * A_TILE;
#pragma unroll
    for (int t = 0; t < BPT; ++t) {
      const int idx = tid + t * 256;
      if (idx < B_ITEMS) {
        const int kl = idx % NT, pj = idx / NT;
        const int col = k0 + kl;
        if (!plain && col < a.k && col < a.tin.act_cols) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int64_t row = r0 + 4 * pj + e;
            if (row < r_end) {
              const int64_t geom = a.tin.escale != nullptr ? geom_of(row, a.rows_per_geom) : 0;
              float zz[CJ];
#pragma unroll
              for (int c = 0; c < CJ; ++c) zz[c] = zv[t][c][e];
              float m;
              const float s = in_scale(a.tin, seed, row, geom, col, m);
              jet_act_fwd<CJ>(a.tin.act, s, zz);
#pragma unroll
              for (int c = 0; c < CJ; ++c) zv[t][c][e] = zz[c];
            }
          }
        }
#pragma unroll
        for (int c = 0; c < CJ; ++c) {
          float4 hi, lo;
          tc::split4(zv[t][c], hi, lo);
          const int j = (c * BR) / 4 + pj;             // chunk of entries c*BR + 4*pj .. +3
          const uint32_t off = j * LBO_B + (kl >> 3) * SBO + (kl & 7) * 16;
          *reinterpret_cast<float4*>(bt + off) = hi;
          *reinterpret_cast<float4*>(bt + B_TILE + off) = lo;
        }
      }
    }
  };

  const int nsteps = (int)((r_end - r_begin + BR - 1) / BR);
  auto process = [&](int i, float (&gv)[APT][4], float (&zv)[BPT][CJ][4]) {
    const int s = i & 1;
    const int64_t r0 = r_begin + (int64_t)i * BR;
    uint8_t* st = smem + s * STAGE_BYTES;
    if (i >= 2) bounded_wait(&mma_done[s], ((i >> 1) - 1) & 1);
    stage_chunk(r0, st, gv, zv);
    if (i + 2 < nsteps) load_chunk(r0 + 2 * BR, gv, zv);
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      const uint32_t abase = tc::smem_u32(st);
      const uint32_t bbase = abase + 2 * A_TILE;
#pragma unroll
      for (int eg = 0; eg < E / 8; ++eg) {
        const uint64_t da_hi = tc::make_smem_desc(abase + eg * 2 * LBO_A, LBO_A, SBO);
        const uint64_t da_lo = tc::make_smem_desc(abase + A_TILE + eg * 2 * LBO_A, LBO_A, SBO);
        const uint64_t db_hi = tc::make_smem_desc(bbase + eg * 2 * LBO_B, LBO_B, SBO);
        const uint64_t db_lo = tc::make_smem_desc(bbase + B_TILE + eg * 2 * LBO_B, LBO_B, SBO);
        tc::mma_tf32(tmem_base, da_hi, db_hi, IDESC, (i > 0 || eg > 0) ? 1u : 0u);
        tc::mma_tf32(tmem_base, da_lo, db_hi, IDESC, 1u);
        tc::mma_tf32(tmem_base, da_hi, db_lo, IDESC, 1u);
      }
      tc::mma_commit(&mma_done[s]);
      if (i + 1 == nsteps) tc::mma_commit(&acc_done);
    }
  };
  if (nsteps > 0) load_chunk(r_begin, gvA, zvA);
  if (nsteps > 1) load_chunk(r_begin + BR, gvB, zvB);
  for (int i = 0; i < nsteps; i += 2) {
    process(i, gvA, zvA);
    if (i + 1 < nsteps) process(i + 1, gvB, zvB);
  }
  float* dst = a.partial + (int64_t)blockIdx.y * a.n * a.k;
  const int q = warp & 3, hcol = warp >> 2;
  const int nn = n0 + 32 * q + lane;
  if (nsteps > 0) {
    bounded_wait(&acc_done, 0);
    tc::tc_fence_after();
  }
#pragma unroll 1
  for (int cb = 0; cb < NT / 32; ++cb) {
    const int col = hcol * (NT / 2) + cb * 16;
    float v[16];
    if (nsteps > 0) {
      tc::tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + col, v);
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] = 0.0f;
    }
    if (nn < a.n) {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int kk = k0 + col + e;
        if (kk < a.k) dst[(int64_t)nn * a.k + kk] = v[e];
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

struct TcDwPlan { int br, nt, tiles_n, tiles_k, splits; int64_t rows_per_split; int64_t chunks, rows_per_chunk; };

static TcDwPlan plan_dw_tc(int cj, int64_t rows, int64_t rows_per_geom, int k, int n) {
  TcDwPlan p;
  p.br = cj == 1 ? 32 : 8;
  const int max_nt = cj == 7 ? 64 : (cj == 5 ? 128 : 256);
  p.nt = k <= 32 ? 32 : (k <= 64 ? 64 : (k <= 128 ? 128 : 256));
  if (p.nt > max_nt) p.nt = max_nt;
  p.tiles_n = (n + 127) / 128;
  p.tiles_k = (k + p.nt - 1) / p.nt;
  const int tiles = p.tiles_n * p.tiles_k;
  int64_t splits = (2 * 148 + tiles - 1) / tiles;
  const int64_t min_rows = cj == 1 ? 256 : 128;              // at least a few pipeline stages per CTA
  const int64_t max_splits = (rows + min_rows - 1) / min_rows;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t rps = (rows + splits - 1) / splits;
  rps = (rps + p.br - 1) / p.br * p.br;
  p.rows_per_split = rps;
  p.splits = (int)((rows + rps - 1) / rps);
  p.rows_per_chunk = rows_per_geom > 0 ? rows_per_geom : 2048;
  p.chunks = (rows + p.rows_per_chunk - 1) / p.rows_per_chunk;
  return p;
}

template <int CJ, int BR, int NT>
static int launch_dw_tc(const TcDwArgs& a, const TcDwPlan& p, cudaStream_t st) {
  constexpr int E = CJ * BR;
  constexpr int SMEM = 2 * (2 * E * 128 * 4 + 2 * E * NT * 4);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(jet_dw_tc_kernel<CJ, BR, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;
    configured = true;
  }
  dim3 grid((unsigned)(p.tiles_n * p.tiles_k), (unsigned)p.splits);
  jet_dw_tc_kernel<CJ, BR, NT><<<grid, 256, SMEM, st>>>(a);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

}  // namespace pcfd

using namespace pcfd;

extern "C" int pcfd_tc_supported_bwd(int32_t cj, int64_t rows, int32_t k, int32_t n) {
  return valid_cj(cj) && rows >= 512 && (int64_t)k * n >= 64;
}

extern "C" int pcfd_tc_jet_linear_bwd_dx(const float* gzout, int64_t gzout_ps, int32_t ldgzout, const float* w,
                                         int32_t ldw, const float* zin, int64_t zin_ps, int32_t ldzin,
                                         const pcfd_intrans_t* tin, float* gzin, int64_t gzin_ps, int32_t ldgzin,
                                         float* gescale, int32_t ldgescale, int32_t cj, int64_t rows,
                                         int64_t rows_per_geom, int32_t k, int32_t n, void* stream) {
  TcDxArgs a{gzout, gzout_ps, ldgzout, w, ldw, zin, zin_ps, ldzin, make_intrans(tin, k), gzin, gzin_ps, ldgzin,
             gescale, ldgescale, rows, rows_per_geom, k, n, 0, 0, 0, 0};
  a.vec_g = al16(gzout) && ldgzout % 4 == 0 && gzout_ps % 4 == 0;
  a.vec_w = al16(w) && ldw % 4 == 0;
  a.vec_z = al16(zin) && ldzin % 4 == 0 && zin_ps % 4 == 0;
  a.vec_out = al16(gzin) && ldgzin % 4 == 0 && gzin_ps % 4 == 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (cj) {
    case 1: return k > 128 ? launch_dx_tc<1, 16, 256, 16>(a, st) : (k > 32 ? launch_dx_tc<1, 16, 128, 16>(a, st) : launch_dx_tc<1, 16, 32, 16>(a, st));
    case 3: return k > 32 ? launch_dx_tc<3, 16, 128, 16>(a, st) : launch_dx_tc<3, 16, 32, 16>(a, st);
    case 4: return k > 32 ? launch_dx_tc<4, 16, 128, 16>(a, st) : launch_dx_tc<4, 16, 32, 16>(a, st);
    case 5: return k > 32 ? launch_dx_tc<5, 16, 64, 8>(a, st) : launch_dx_tc<5, 16, 32, 8>(a, st);
    case 7: return k > 32 ? launch_dx_tc<7, 8, 64, 8>(a, st) : launch_dx_tc<7, 8, 32, 8>(a, st);
  }
  return PCFD_ERR_ARG;
}

extern "C" size_t pcfd_tc_dw_workspace_bytes(int32_t cj, int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n) {
  TcDwPlan p = plan_dw_tc(cj, rows, rows_per_geom, k, n);
  return ((size_t)p.splits * n * k + (size_t)p.chunks * ((p.rows_per_chunk + 127) / 128) * n) * sizeof(float) + 256;
}

// writes partial[splits][n][k] at the start of `workspace`; returns the number of splits through *splits_out
extern "C" int pcfd_tc_jet_linear_bwd_dw_partials(const float* gzout, int64_t gzout_ps, int32_t ldgzout,
                                                  const float* zin, int64_t zin_ps, int32_t ldzin,
                                                  const pcfd_intrans_t* tin, int32_t cj, int64_t rows,
                                                  int64_t rows_per_geom, int32_t k, int32_t n, void* workspace,
                                                  int* splits_out, void* stream) {
  TcDwPlan p = plan_dw_tc(cj, rows, rows_per_geom, k, n);
  TcDwArgs a{gzout, gzout_ps, ldgzout, zin, zin_ps, ldzin, make_intrans(tin, k), reinterpret_cast<float*>(workspace),
             rows, rows_per_geom, p.rows_per_split, k, n, p.tiles_k, 0, 0};
  a.vec_g = al16(gzout) && ldgzout % 4 == 0 && gzout_ps % 4 == 0;
  a.vec_z = al16(zin) && ldzin % 4 == 0 && zin_ps % 4 == 0;
  *splits_out = p.splits;
  cudaStream_t st = (cudaStream_t)stream;
#define PCFD_DW_CASE(CJ_, BR_)                                                 \
  if (p.nt == 32) return launch_dw_tc<CJ_, BR_, 32>(a, p, st);                 \
  if (p.nt == 64) return launch_dw_tc<CJ_, BR_, 64>(a, p, st);                 \
  if (p.nt == 128) return launch_dw_tc<CJ_, BR_, 128>(a, p, st);               \
  return launch_dw_tc<CJ_, BR_, 256>(a, p, st);
  switch (cj) {
    case 1: PCFD_DW_CASE(1, 32)
    case 3: PCFD_DW_CASE(3, 8)
    case 4: PCFD_DW_CASE(4, 8)
    case 5: if (p.nt == 32) return launch_dw_tc<5, 8, 32>(a, p, st); if (p.nt == 64) return launch_dw_tc<5, 8, 64>(a, p, st); return launch_dw_tc<5, 8, 128>(a, p, st);
    case 7: if (p.nt == 32) return launch_dw_tc<7, 8, 32>(a, p, st); return launch_dw_tc<7, 8, 64>(a, p, st);
  }
#undef PCFD_DW_CASE
  return PCFD_ERR_ARG;
}
