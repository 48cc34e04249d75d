// Probe for the building blocks of the warp-specialised tcgen05 engine (run on a B200):
//   1. TMA (cuTensorMapEncodeTiled through the driver entry point) into 64B / 128B swizzled tiles,
//      consumed by tcgen05.mma kind::tf32 with K-major swizzled descriptors;
//   2. MN-major 128B-swizzled operands (the natural layout of the dW contraction over rows);
//   3. what the tensor core does with the low 13 mantissa bits of an fp32 operand word
//      (truncate vs round) -- decides whether raw fp32 tiles can serve as the "hi" TF32 operand;
//   4. TMA store of a swizzled staging tile.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/probe/ws_probe scripts/probe/ws_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../porous-cfd_b200/csrc/ws_common.cuh"
using namespace pcfd;

constexpr int M = 128, N = 128;

template <int SWB, bool MN>
__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmA,
                                                const __grid_constant__ CUtensorMap tmB,
                                                const __grid_constant__ CUtensorMap tmD, float* out, int K,
                                                uint32_t mn_group_stride) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full, done;
  __shared__ uint32_t tbase;
  constexpr int KC = MN ? 16 : SWB / 4;                 // contraction entries per chunk
  constexpr int TILE = MN ? 4 * KC * 128 : 128 * SWB;   // bytes of one operand chunk
  uint8_t* As = smem;
  uint8_t* Bs = smem + TILE;
  uint8_t* Ds = smem + 2 * TILE;                        // 4 warps x [32 rows x 128 B] staging
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tc::tmem_alloc(&tbase, 128);
  if (tid == 0) {
    tc::mbar_init(&full, 1);
    tc::mbar_init(&done, 1);
    tc::fence_mbar_init();
    ws::prefetch_tmap(&tmA);
    ws::prefetch_tmap(&tmB);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tbase;
  const uint32_t idesc = tc::make_idesc_tf32(128, N, MN, MN);

  if (tid == 0) {
    uint32_t ph = 0;
    for (int k0 = 0; k0 < K; k0 += KC) {
      if (k0 > 0) { tc::bounded_wait(&done, ph ^ 1); }   // previous chunk's MMAs finished reading smem
      ws::mbar_expect_tx(&full, 2 * TILE);
      if (MN) {
        ws::tma_load_3d(As, &tmA, 0, k0, 0, &full);
        ws::tma_load_3d(Bs, &tmB, 0, k0, 0, &full);
      } else {
        ws::tma_load_2d(As, &tmA, k0, 0, &full);
        ws::tma_load_2d(Bs, &tmB, k0, 0, &full);
      }
      tc::bounded_wait(&full, ph);
      tc::tc_fence_after();
      for (int ks = 0; ks < KC / 8; ++ks) {
        uint64_t da, db;
        if (MN) {
          da = ws::desc_mnmajor(tc::smem_u32(As) + ks * 1024, KC * 128, mn_group_stride);
          db = ws::desc_mnmajor(tc::smem_u32(Bs) + ks * 1024, KC * 128, mn_group_stride);
        } else {
          da = ws::desc_kmajor<SWB>(tc::smem_u32(As) + ks * 32);
          db = ws::desc_kmajor<SWB>(tc::smem_u32(Bs) + ks * 32);
        }
        tc::mma_tf32(tmem, da, db, idesc, (k0 > 0 || ks > 0) ? 1u : 0u);
      }
      tc::mma_commit(&done);
      ph ^= 1;
    }
    tc::bounded_wait(&done, ph ^ 1);
  }
  __syncthreads();
  tc::tc_fence_after();
  // direct store + staged TMA store of each 32-column block
  for (int cb = 0; cb < N / 32; ++cb) {
    uint32_t r[32];
    ws::tmem_ld32_nowait(tmem + ((uint32_t)(32 * warp) << 16) + cb * 32, r);
    ws::tmem_ld_wait();
    for (int e = 0; e < 32; ++e) out[(32 * warp + lane) * N + cb * 32 + e] = __uint_as_float(r[e]);
    uint8_t* st = Ds + warp * 4096;
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<uint4*>(st + ws::swz<128>(lane, j)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    tc::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      ws::tma_store_2d(&tmD, st, cb * 32, 32 * warp);
      ws::tma_store_commit();
      ws::tma_store_wait_read<0>();
    }
    __syncwarp();
  }
  if (lane == 0) ws::tma_store_wait_all<0>();
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

static float trunc_h(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float rna_h(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

template <int SWB, bool MN>
static int run(int K, int lda, const char* name, uint32_t gs = 512) {
  // K-major: A[m][k] (ld = lda >= K); MN-major: A[k][m] (ld = lda >= 128)
  const size_t na = MN ? (size_t)K * lda : (size_t)M * lda;
  std::vector<float> hA(na), hB(na), hD(M * N), hD2(M * N);
  srand(7);
  for (auto& v : hA) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : hB) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD, *dD2;
  cudaMalloc(&dA, na * 4); cudaMalloc(&dB, na * 4); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dD2, M * N * 4);
  cudaMemcpy(dA, hA.data(), na * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), na * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, M * N * 4); cudaMemset(dD2, 0, M * N * 4);
  CUtensorMap tmA, tmB, tmD;
  bool ok = true;
  if (MN) {
    uint64_t dims[3] = {32, (uint64_t)K, 4};
    uint64_t str[2] = {(uint64_t)lda * 4, 128};
    uint32_t box[3] = {32, 16, 4};
    ok &= ws::make_tmap(&tmA, dA, 3, dims, str, box, ws::SW128_ATOM32);
    ok &= ws::make_tmap(&tmB, dB, 3, dims, str, box, ws::SW128_ATOM32);
  } else {
    uint64_t dims[2] = {(uint64_t)K, 128};
    uint64_t str[1] = {(uint64_t)lda * 4};
    uint32_t box[2] = {SWB / 4, 128};
    ok &= ws::make_tmap(&tmA, dA, 2, dims, str, box, SWB);
    ok &= ws::make_tmap(&tmB, dB, 2, dims, str, box, SWB);
  }
  {
    uint64_t dims[2] = {128, 128};
    uint64_t str[1] = {128 * 4};
    uint32_t box[2] = {32, 32};
    ok &= ws::make_tmap(&tmD, dD2, 2, dims, str, box, 128);
  }
  if (!ok) { printf("%s: tensor map encode FAILED\n", name); return 1; }
  const int smem = 1024 + 2 * 65536 / 2 + 16384 + 1024;
  cudaFuncSetAttribute(probe<SWB, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<SWB, MN><<<1, 128, smem>>>(tmA, tmB, tmD, dD, K, gs);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return 2; }
  cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hD2.data(), dD2, M * N * 4, cudaMemcpyDeviceToHost);
  double et = 0, er = 0, ef = 0, es = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double st = 0, sr = 0, sf = 0;
      for (int k = 0; k < K; ++k) {
        const float a = MN ? hA[(size_t)k * lda + m] : hA[(size_t)m * lda + k];
        const float b = MN ? hB[(size_t)k * lda + n] : hB[(size_t)n * lda + k];
        st += (double)trunc_h(a) * trunc_h(b);
        sr += (double)rna_h(a) * rna_h(b);
        sf += (double)a * b;
      }
      const double d = hD[m * N + n];
      et = fmax(et, fabs(d - st)); er = fmax(er, fabs(d - sr)); ef = fmax(ef, fabs(d - sf));
      es = fmax(es, fabs((double)hD2[m * N + n] - d));
    }
  printf("%-28s K=%3d lda=%3d : max|D-trunc|=%.3e  max|D-rna|=%.3e  max|D-fp32|=%.3e  |TMAstore-direct|=%.3e  %s\n", name, K,
         lda, et, er, ef, es, (et < 1e-4 || er < 1e-4) ? "OK" : "MISMATCH");
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dD2);
  return 0;
}

int main() {
  int rc = 0;
  rc |= run<128, false>(32, 32, "K-major SW128");
  rc |= run<128, false>(96, 100, "K-major SW128");
  rc |= run<64, false>(16, 16, "K-major SW64");
  rc |= run<64, false>(80, 96, "K-major SW64");
  rc |= run<128, true>(16, 128, "MN-major SW128/32B gs=512", 512);
  rc |= run<128, true>(64, 160, "MN-major SW128/32B gs=512", 512);
  rc |= run<128, true>(16, 128, "MN-major SW128/32B gs=1024", 1024);
  rc |= run<128, true>(16, 128, "MN-major SW128/32B gs=256", 256);
  return rc;
}
