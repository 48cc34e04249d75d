// tcgen05.mma kind::tf32 issue-rate probe (run on a B200): cycles per MMA for M = 128, N in {64, 128, 256},
// K = 8, with K-major / MN-major operand descriptors, accumulating into one or several TMEM tiles.
// Shared-memory contents are whatever is there (only timing matters).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/probe/mma_rate scripts/probe/mma_rate.cu
#include <cstdio>
#include <cstdlib>

#include "../../porous-cfd_b200/csrc/ws_common.cuh"
using namespace pcfd;

template <int N, bool AMN, bool BMN>
__global__ void __launch_bounds__(128, 1) rate(long long* out, int iters, int ntiles, int distinct_ops) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (warp == 0) tc::tmem_alloc(&tbase, 512);
  if (tid == 0) { tc::mbar_init(&done, 1); tc::fence_mbar_init(); }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tbase;
  constexpr uint32_t idesc = tc::make_idesc_tf32(128, N, AMN, BMN);
  if (tid == 0) {
    const uint32_t sa = tc::smem_u32(smem), sb = sa + 64 * 1024;
    long long t0 = clock64();
    uint64_t da[2], db[2];
    for (int o = 0; o < 2; ++o) {
      da[o] = AMN ? ws::desc_mnmajor(sa + o * 1024, 16 * 128, 512) : ws::desc_kmajor<64>(sa + o * 32);
      db[o] = BMN ? ws::desc_mnmajor(sb + o * 1024, 16 * 128, 512) : ws::desc_kmajor<64>(sb + o * 32);
    }
    if (distinct_ops == 2) {
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) tc::mma_tf32(tmem + (ntiles > 1 ? (j % 2) * N : 0), da[j & 1], db[j & 1], idesc, 1u);
      }
    } else {
      // commit after every MMA (as a per-stage pipeline with one MMA per stage would)
      for (int i = 0; i < iters; ++i) {
        tc::mma_tf32(tmem, da[i & 1], db[i & 1], idesc, 1u);
        if ((i & 3) == 3) tc::mma_commit(&done);
      }
      for (int i = 0; i < iters / 4; ++i) tc::bounded_wait(&done, i & 1);
      tc::mbar_init(&done, 1);
    }
    tc::mma_commit(&done);
    tc::bounded_wait(&done, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// distinct operand tiles every MMA: 8 stages x (BK / 8) K steps, A and B walking through 128 KB of shared memory
// LA / LB: 0 = K-major SW64 (16 entries per row), 1 = K-major SW128 (32 entries per row), 2 = MN-major 32B-atom
template <int N, int LA, int LB>
__global__ void __launch_bounds__(128, 1) rate2(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, warp = ws::uniform_warp_id();
  for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (warp == 0) tc::tmem_alloc(&tbase, 512);
  if (tid == 0) { tc::mbar_init(&done, 1); tc::fence_mbar_init(); }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tbase;
  constexpr uint32_t idesc = tc::make_idesc_tf32(128, N, LA == 2, LB == 2);
  if (warp == 0) {
    const uint32_t sa = tc::smem_u32(smem), sb = sa + 100 * 1024;
    const uint64_t abase = LA == 2 ? ws::desc_mnmajor(sa, 16 * 128, 512) : (LA == 1 ? ws::desc_kmajor<128>(sa) : ws::desc_kmajor<64>(sa));
    const uint64_t bbase = LB == 2 ? ws::desc_mnmajor(sb, 16 * 128, 512) : (LB == 1 ? ws::desc_kmajor<128>(sb) : ws::desc_kmajor<64>(sb));
    constexpr int KS_A = LA == 1 ? 4 : 2, KS_B = LB == 1 ? 4 : 2;     // K steps per tile row block
    constexpr int STEP_A = LA == 2 ? 1024 : 32, STEP_B = LB == 2 ? 1024 : 32;
    constexpr int TILE_A = LA == 1 ? 128 * 128 : (LA == 0 ? 128 * 64 : 4 * 16 * 128);
    constexpr int TILE_B = LB == 1 ? N * 128 : (LB == 0 ? N * 64 : (N / 32) * 16 * 128);
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 16) {
      if (ws::elect_one()) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int ta = (j / KS_A) % (96 * 1024 / TILE_A), tb = (j / KS_B) % (96 * 1024 / TILE_B);
          tc::mma_tf32(tmem, abase + ((ta * TILE_A + (j % KS_A) * STEP_A) >> 4), bbase + ((tb * TILE_B + (j % KS_B) * STEP_B) >> 4), idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (ws::elect_one()) tc::mma_commit(&done);
    __syncwarp();
    tc::bounded_wait(&done, 0);
    long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

template <int N, int LA, int LB>
static void run2(const char* name) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const int smem = 210 * 1024, grid = 148, iters = 4096;
  cudaFuncSetAttribute(rate2<N, LA, LB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate2<N, LA, LB><<<grid, 128, smem>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-40s N=%3d : %7.1f clk/MMA  (floor %5.1f)  %s\n", name, N, (double)mx / iters, 128.0 * N / 256.0,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

template <int N, bool AMN, bool BMN>
static void run(const char* name, int ntiles, int grid) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(rate<N, AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4096;
  rate<N, AMN, BMN><<<grid, 128, smem>>>(d, iters, ntiles, 2);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double cyc = (double)mx / iters;
  printf("%-34s N=%3d tiles=%d grid=%3d : %7.1f clk/MMA  (ideal %5.1f)  %s\n", name, N, ntiles, grid, cyc,
         128.0 * N * 8 / 2048.0 / 2.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run2<128, 0, 0>("A K SW64,  B K SW64");
  run2<128, 1, 1>("A K SW128, B K SW128");
  run2<128, 0, 2>("A K SW64,  B MN");
  run2<128, 1, 2>("A K SW128, B MN");
  run2<128, 2, 2>("A MN, B MN");
  run2<64, 0, 0>("A K SW64,  B K SW64");
  run2<64, 1, 1>("A K SW128, B K SW128");
  run2<64, 2, 2>("A MN, B MN");
  run2<256, 0, 0>("A K SW64,  B K SW64");
  run2<256, 1, 1>("A K SW128, B K SW128");
  run2<256, 2, 2>("A MN, B MN");
  run2<256, 1, 2>("A K SW128, B MN");
  return 0;

  for (int grid : {1, 148}) {
    run<64, false, false>("A K-major,  B K-major", 1, grid);
    run<64, false, true>("A K-major,  B MN-major", 1, grid);
    run<64, true, false>("A MN-major, B K-major", 1, grid);
    run<64, true, true>("A MN-major, B MN-major", 1, grid);
    run<128, false, false>("A K-major,  B K-major", 1, grid);
    run<128, false, true>("A K-major,  B MN-major", 1, grid);
    run<128, true, false>("A MN-major, B K-major", 1, grid);
    run<128, true, true>("A MN-major, B MN-major", 1, grid);
    run<256, false, false>("A K-major,  B K-major", 1, grid);
    run<256, false, true>("A K-major,  B MN-major", 1, grid);
    run<256, true, false>("A MN-major, B K-major", 1, grid);
    run<256, true, true>("A MN-major, B MN-major", 1, grid);
    run<128, true, true>("A MN-major, B MN-major 2 tiles", 2, grid);
    run<128, false, false>("A K-major,  B K-major 2 tiles", 2, grid);
    run<256, true, true>("A MN-major, B MN-major 2 tiles", 2, grid);
  }
  return 0;
}
