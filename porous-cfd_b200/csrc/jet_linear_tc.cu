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

__device__ __forceinline__ void bounded_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = tc::smem_u32(bar);
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();   // a wedged tensor pipe must fail the launch, not hang the GPU
}

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
  const int64_t geom = (a.rows_per_geom > 0 && valid) ? row / a.rows_per_geom : 0;
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
  float z[CPT][CJ][4];
  float4 wv[BPT];

  auto load_chunk = [&](int i) {
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

  auto stage_chunk = [&](int i, uint8_t* st) {
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

  load_chunk(0);
  for (int i = 0; i < nchunks; ++i) {
    const int s = i & 1;
    uint8_t* st = smem + s * STAGE_BYTES;
    if (i >= 2) bounded_wait(&mma_done[s], ((i >> 1) - 1) & 1);     // MMAs that read this stage are done
    stage_chunk(i, st);
    if (i + 1 < nchunks) load_chunk(i + 1);                          // HBM loads in flight across the barrier
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
  }

  bounded_wait(&acc_done, 0);
  tc::tc_fence_after();

  // epilogue: warp (q = warp & 3) owns TMEM lanes [32q, 32q+32); the two warpgroups split the columns
  const int q = warp & 3, hcol = warp >> 2;
  const int64_t orow = row0 + 32 * q + lane;
  const bool ovalid = orow < a.rows;
  const int64_t ogeom = (a.rows_per_geom > 0 && ovalid) ? orow / a.rows_per_geom : 0;
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
  return valid_cj(cj) && rows >= 512 && k >= 16 && n >= 16;
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
    case 1: return n > 128 ? launch_fwd_tc<1, 16, 256>(a, st) : launch_fwd_tc<1, 16, 128>(a, st);
    case 3: return launch_fwd_tc<3, 16, 128>(a, st);
    case 4: return launch_fwd_tc<4, 16, 128>(a, st);
    case 5: return launch_fwd_tc<5, 16, 64>(a, st);
    case 7: return launch_fwd_tc<7, 8, 64>(a, st);
  }
  return PCFD_ERR_ARG;
}
