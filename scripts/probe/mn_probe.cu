// Probe: which shared-memory word does tcgen05.mma (kind::tf32, SWIZZLE_NONE) read as B[k][n] for an
// MN-major B descriptor with given LBO / SBO?  A = [I8 | 0] so D[m][n] = B[k=m][n] for m < 8.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../porous-cfd_b200/csrc/tc_common.cuh"
using namespace pcfd;

__global__ void probe(uint32_t lbo, uint32_t sbo, int b_mn, int a_mn, float* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tbase;
  float* A = reinterpret_cast<float*>(smem);            // 128 x 8 K-major no-swizzle: chunk(row, j) at j*2048 + (row/8)*128 + (row%8)*16
  float* B = reinterpret_cast<float*>(smem + 8192);     // 2048 words, value = word index
  const int tid = threadIdx.x;
  for (int i = tid; i < 2048; i += blockDim.x) { A[i] = 0.f; }
  for (int i = tid; i < 2048; i += blockDim.x) { B[i] = (float)i; }
  __syncthreads();
  if (!a_mn) {
    if (tid < 8) {   // A[row=tid][k=tid] = 1 : chunk j = k/4, word k%4
      int row = tid, k = tid;
      A[((k / 4) * 2048 + (row / 8) * 128 + (row % 8) * 16) / 4 + (k % 4)] = 1.f;
    }
  } else {
    if (tid < 8) {   // MN-major A (M=128,K=8): chunk(k, g=m/4) at g*128 + k*16 (+ (m%4)*4)
      int m = tid, k = tid;
      A[((m / 4) * 128 + k * 16) / 4 + (m % 4)] = 1.f;
    }
  }
  if (tid < 32) tc::tmem_alloc(&tbase, 64);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (tid == 0) {
    uint64_t da = a_mn ? tc::make_smem_desc(tc::smem_u32(A), 4096, 128) : tc::make_smem_desc(tc::smem_u32(A), 2048, 128);
    uint64_t db = tc::make_smem_desc(tc::smem_u32(B), lbo, sbo);
    uint32_t idesc = tc::make_idesc_tf32(128, 64, a_mn != 0, b_mn != 0);
    tc::mma_tf32(tbase, da, db, idesc, 0u);
    tc::mma_commit(&bar);
  }
  tc::bounded_wait(&bar, 0);
  tc::tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  if (warp < 4) {
    for (int cb = 0; cb < 4; ++cb) {
      float v[16];
      tc::tmem_ld16(tbase + ((uint32_t)(32 * warp) << 16) + cb * 16, v);
      for (int e = 0; e < 16; ++e) out[(32 * warp + lane) * 64 + cb * 16 + e] = v[e];
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tbase, 64);
}

int main() {
  float* out; cudaMalloc(&out, 128 * 64 * 4);
  float h[128 * 64];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  struct { uint32_t lbo, sbo; int b_mn, a_mn; } cfg[] = {{1024, 128, 0, 0}, {2048, 128, 1, 0}, {128, 2048, 1, 0}, {128, 1024, 1, 0},
                                                          {1024, 128, 1, 0}, {256, 128, 1, 0}, {128, 256, 1, 0}, {2048, 128, 1, 1}, {128, 2048, 0, 1}};
  for (auto& c : cfg) {
    cudaMemset(out, 0, sizeof(h));
    probe<<<1, 128, 32768>>>(c.lbo, c.sbo, c.b_mn, c.a_mn, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("== lbo=%u sbo=%u b_mn=%d a_mn=%d : %s\n", c.lbo, c.sbo, c.b_mn, c.a_mn, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    for (int m = 0; m < 8; ++m) {
      printf("k=%d:", m);
      for (int n = 0; n < 12; ++n) printf(" %4d", (int)h[m * 64 + n]);
      printf(" ... n=32: %4d %4d  n=63: %4d\n", (int)h[m * 64 + 32], (int)h[m * 64 + 33], (int)h[m * 64 + 63]);
    }
  }
  return 0;
}
