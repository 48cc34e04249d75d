// TMA / swizzled-operand building blocks of the warp-specialised tcgen05 jet engine (engine 2).
//
//   * host: CUtensorMap encoding through the driver entry point (no link-time libcuda dependency),
//   * device: cp.async.bulk.tensor loads/stores, mbarrier transaction counts, swizzled UMMA
//     shared-memory descriptors (K-major and MN-major, 64-byte and 128-byte swizzle).
//
// Swizzled tiles: a tile is a stack of "rows" of SWB bytes (SWB = 64 or 128); the 16-byte chunk j
// of row r is stored at  r*SWB + ((j ^ f(r)) * 16)  with f(r) = r % 8 (SWB = 128) or (r / 2) % 4
// (SWB = 64) -- the address-bit XOR Swizzle<3,4,3> / Swizzle<2,4,3> of the CUTLASS canonical
// layouts, which is also what TMA writes for CU_TENSOR_MAP_SWIZZLE_128B / _64B into a tile whose
// base is 1024-byte aligned.
//   K-major operand  : row = M/N index, the SWB bytes of a row are consecutive K entries.
//                      descriptor: SBO = 8 rows * SWB, LBO unused (1); the K step of one MMA
//                      (8 tf32 = 32 bytes) advances the start address by 32 bytes.
//   MN-major operand : fp32 operands transposed by the tensor core need the 32-byte-atom variant
//                      (measured: plain SWIZZLE_128B MN-major gives wrong products for kind::tf32):
//                      row = K index, a row holds 32 consecutive M/N entries, 32-byte unit u of row
//                      r is stored at r*128 + ((u ^ (r % 4)) * 32); blocks of 32 M/N entries are
//                      LBO bytes apart, 4-row K groups SBO bytes apart.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <utility>

#include "common.cuh"
#include "tc_common.cuh"

namespace pcfd {
namespace ws {

// geometry of the row-tile kernels (forward, dX): a stage holds 16 contraction entries of 8 slabs of 32 rows
constexpr int BK = 16;                 // contraction entries per stage
constexpr int SLAB_BYTES = 32 * 64;    // one slab (32 rows) of a stage, 64-byte rows
constexpr int A_BYTES = 8 * SLAB_BYTES;
constexpr int STAGES = 4;

// swizzle selector of the MN-major fp32 layout: 32-byte units XORed within a 128-byte row
// (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B / UMMA layout type 1, "SWIZZLE_128B_BASE32B")
constexpr int SW128_ATOM32 = 132;

// ---- host: tensor maps --------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp32 tensor of `rank` dims (dim 0 innermost, contiguous); strides[i] = byte stride of dim i+1.
// Out-of-bounds elements of a box read as zero.  swizzle_bytes: 0, 64 or 128.
inline bool make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return false;
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == SW128_ATOM32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// Grid of a persistent kernel that walks `total` work items with stride gridDim.x: the smallest grid that keeps the same
// number of items per CTA as one CTA per SM would (the kernel ends when the busiest CTA ends, so the extra CTAs of a
// ragged last wave buy nothing), which leaves the remaining SMs to the kernels of the step's other streams.
// 750 row tiles on 148 SMs: 6 items on the busiest CTA either way -> 125 CTAs.
inline int balanced_grid(int total) {
  const int sms = num_sms();
  if (total <= sms) return total;
  const int per_cta = (total + sms - 1) / sms;
  return (total + per_cta - 1) / per_cta;
}

// ---- device: TMA ----------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(tc::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(tc::smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, int x, int y, int z, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(tc::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(z),
        "r"(tc::smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tm, int x, int y, int z, int w,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(tc::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(z), "r"(w),
        "r"(tc::smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* smem_src, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(tc::smem_u32(smem_src)), "r"(x), "r"(y), "r"(z)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(tc::smem_u32(smem_src)), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the shared-memory source of all but the newest `N` store groups may be overwritten
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// ---- device: swizzle ----------------------------------------------------------------------------
// byte offset of 16-byte chunk j of row r inside a swizzled tile with SWB-byte rows
template <int SWB>
__device__ __forceinline__ uint32_t swz(uint32_t r, uint32_t j) {
  if (SWB == 128) return r * 128u + ((j ^ (r & 7u)) << 4);
  if (SWB == 64) return r * 64u + ((j ^ ((r >> 1) & 3u)) << 4);
  return r * (uint32_t)SWB + (j << 4);
}

// ---- device: descriptors ------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_desc_sw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, int swb) {
  const uint64_t layout = swb == 128 ? 2ull : (swb == 64 ? 4ull : (swb == 32 ? 6ull : (swb == SW128_ATOM32 ? 1ull : 0ull)));
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}
// K-major tile of SWB-byte rows starting at smem_addr (+ 32 bytes per K step of 8)
template <int SWB>
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t smem_addr) { return make_desc_sw(smem_addr, 16, 8 * SWB, SWB); }
// MN-major fp32 tile (rows = K entries, 128 bytes = 32 M/N entries each, 32-byte units XORed with row % 4):
// blocks of 32 M/N entries `block_stride` bytes apart, K groups of 4 rows `group_stride` bytes apart
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t smem_addr, uint32_t block_stride, uint32_t group_stride = 512) {
  return make_desc_sw(smem_addr, block_stride, group_stride, SW128_ATOM32);
}

// 32 consecutive accumulator columns of this thread's lane -> registers (no wait)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One lane of the (converged) warp.  The producer / MMA-issuer warps run their loops with all 32 lanes so that
// addresses, descriptors and coordinates stay in uniform registers (UTMALDG / UTCHMMA take uniform operands; code
// under `if (lane == 0)` makes the compiler wrap every such instruction in an ELECT + R2UR "waterfall" loop, which
// costs ~150 cycles per MMA -- measured with scripts/probe/mma_rate.cu), and only the issue itself is predicated.
// elect.sync returns the same leader for the same member mask, so tcgen05.commit sees the MMAs of the same thread.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// warp index as a provably warp-uniform value
__device__ __forceinline__ int uniform_warp_id() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// named barrier among `nthreads` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// truncation split: hi = the 19 bits the tensor core reads (raw fp32 words are valid TF32 operands),
// lo = x - hi (exact in fp32)
__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// ---- device: A operand in tensor memory (lane = row, one 32-bit column per contraction entry; scripts/probe/ts_probe.cu)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

// ---- device: input transform of a staged chunk ----------------------------------------------------
// activation jet of the (up to) 4 columns of one staged 16-byte chunk, all channels
template <int CJ, int ACT, bool SCALED, bool FULL>
__device__ __forceinline__ void transform_chunk_impl(float (&v)[CJ][4], const InTrans& tin, uint32_t hseed, int64_t row,
                                                     int64_t geom, int col0, int ncols) {
  uint32_t hrow = 0;
  if (SCALED && tin.drop_p > 0.0f) hrow = dropout_row_hash(hseed, row);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (FULL || e < ncols) {
      float sc = 1.0f;
      if (SCALED) {
        if (tin.drop_p > 0.0f) sc = dropout_from_row(hrow, row, col0 + e, tin.drop_p, tin.inv_keep);
        if (tin.escale != nullptr) sc *= __ldg(tin.escale + geom * tin.ldescale + col0 + e);
      }
      float zz[CJ];
#pragma unroll
      for (int c = 0; c < CJ; ++c) zz[c] = v[c][e];
      jet_act_fwd_t<CJ, ACT>(sc, zz);
#pragma unroll
      for (int c = 0; c < CJ; ++c) v[c][e] = zz[c];
    }
  }
}
// all four columns active (the common case) takes the predicate-free path
template <int CJ, int ACT, bool SCALED>
__device__ __forceinline__ void transform_chunk(float (&v)[CJ][4], const InTrans& tin, uint32_t hseed, int64_t row,
                                                int64_t geom, int col0, int ncols) {
  if (ncols >= 4) transform_chunk_impl<CJ, ACT, SCALED, true>(v, tin, hseed, row, geom, col0, ncols);
  else transform_chunk_impl<CJ, ACT, SCALED, false>(v, tin, hseed, row, geom, col0, ncols);
}
template <int CJ>
__device__ __forceinline__ void transform_dispatch(float (&v)[CJ][4], const InTrans& tin, bool scaled, uint32_t hseed,
                                                   int64_t row, int64_t geom, int col0, int ncols) {
  if (tin.act == PCFD_ACT_SILU) {
    if (scaled) transform_chunk<CJ, PCFD_ACT_SILU, true>(v, tin, hseed, row, geom, col0, ncols);
    else transform_chunk<CJ, PCFD_ACT_SILU, false>(v, tin, hseed, row, geom, col0, ncols);
  } else if (tin.act == PCFD_ACT_TANH) {
    if (scaled) transform_chunk<CJ, PCFD_ACT_TANH, true>(v, tin, hseed, row, geom, col0, ncols);
    else transform_chunk<CJ, PCFD_ACT_TANH, false>(v, tin, hseed, row, geom, col0, ncols);
  } else {
    transform_chunk<CJ, PCFD_ACT_NONE, true>(v, tin, hseed, row, geom, col0, ncols);
  }
}

}  // namespace ws
}  // namespace pcfd
