// Probe: tcgen05.mma kind::tf32 with the A operand in tensor memory (".ts" form).  A[128][K] is written to TMEM with
// tcgen05.st.32x32b (lane = row, one 32-bit column per K entry), B[N][K] sits in shared memory as a K-major SW64 tile
// written by ordinary stores.  Checks D = A * B^T against the host (TF32-truncated operands).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/probe/ts_probe scripts/probe/ts_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../porous-cfd_b200/csrc/ws_common.cuh"
using namespace pcfd;

constexpr int M = 128, N = 128, K = 16;

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(const float* A, const float* B, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tc::tmem_alloc(&tbase, 256);
  if (tid == 0) { tc::mbar_init(&done, 1); tc::fence_mbar_init(); }
  // B tile: K-major, 64-byte rows (16 entries), SW64
  for (int i = tid; i < N * 4; i += 128) {
    const int n = i >> 2, j = i & 3;
    *reinterpret_cast<float4*>(smem + ws::swz<64>(n, j)) = *reinterpret_cast<const float4*>(B + n * K + j * 4);
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tbase;
  // A: lane = row (this thread's row = tid), columns 128..143 of the allocation
  uint32_t r[16];
  for (int k = 0; k < 16; ++k) r[k] = __float_as_uint(A[tid * K + k]);
  tmem_st16(tmem + ((uint32_t)(32 * warp) << 16) + 128, r);
  tmem_st_wait();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_tf32(128, N, false, false);
    for (int ks = 0; ks < K / 8; ++ks)
      mma_tf32_ts(tmem, tmem + 128 + ks * 8, ws::desc_kmajor<64>(tc::smem_u32(smem) + ks * 32), idesc, ks > 0 ? 1u : 0u);
    tc::mma_commit(&done);
  }
  tc::bounded_wait(&done, 0);
  tc::tc_fence_after();
  for (int cb = 0; cb < N / 32; ++cb) {
    uint32_t v[32];
    ws::tmem_ld32_nowait(tmem + ((uint32_t)(32 * warp) << 16) + cb * 32, v);
    ws::tmem_ld_wait();
    for (int e = 0; e < 32; ++e) out[(32 * warp + lane) * N + cb * 32 + e] = __uint_as_float(v[e]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

static float trunc_h(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  std::vector<float> hA(M * K), hB(N * K), hD(M * N);
  srand(1);
  for (auto& v : hA) v = (float)rand() / RAND_MAX - 0.5f;
  for (auto& v : hB) v = (float)rand() / RAND_MAX - 0.5f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dD, hD.size() * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  probe<<<1, 128, 32 * 1024>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)trunc_h(hA[m * K + k]) * (double)trunc_h(hB[n * K + k]);
      worst = fmax(worst, fabs(ref - hD[m * N + n]));
    }
  printf("A-in-TMEM tf32 MMA: max |D - ref| = %.3e  %s  (%s)\n", worst, worst < 1e-5 ? "OK" : "MISMATCH", cudaGetErrorString(e));
  return worst < 1e-5 ? 0 : 1;
}
