// Point-set kernels of the PointNet++ set abstraction (farthest point sampling, ball query, edge
// gather / scatter) and the FoamData gathers.  Index results are bit-exact against
// oracle/pyg_restate.py: squared distances use separately rounded fp32 multiplies and adds in
// coordinate order (no FMA contraction), ties resolve to the lowest index.
#include <cstdlib>

#include "common.cuh"

namespace pcfd {

template <int DIMS>
__device__ __forceinline__ float sqdist(const float* __restrict__ p, const float* __restrict__ c) {
  float d0 = __fsub_rn(p[0], c[0]);
  float acc = __fmul_rn(d0, d0);
#pragma unroll
  for (int d = 1; d < DIMS; ++d) {
    float dd = __fsub_rn(p[d], c[d]);
    acc = __fadd_rn(acc, __fmul_rn(dd, dd));
  }
  return acc;
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}

// One CTA per geometry.  Coordinates and running min-distances live in shared memory; one barrier
// per sample: warps publish their best (distance, index) key into a double-buffered slot array and
// every warp re-reduces the slots, so no second barrier is needed to broadcast the winner.
template <int DIMS>
__global__ void __launch_bounds__(1024) fps_kernel(const float* __restrict__ pos, int n, int m,
                                                   int64_t* __restrict__ idx_out) {
  extern __shared__ __align__(16) float smem[];
  float* sp = smem;                 // [n][DIMS]
  float* sd = smem + (size_t)n * DIMS;  // [n]
  __shared__ unsigned long long slot[2][32];
  const int g = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const float* gp = pos + (size_t)g * n * DIMS;
  for (int i = tid; i < n * DIMS; i += nt) sp[i] = gp[i];
  __syncthreads();

  int cur = 0;
  if (tid == 0) idx_out[(size_t)g * m] = (int64_t)g * n;
  for (int s = 1; s < m; ++s) {
    float c[DIMS];
#pragma unroll
    for (int d = 0; d < DIMS; ++d) c[d] = sp[cur * DIMS + d];
    unsigned long long best = 0ULL;
    for (int p = tid; p < n; p += nt) {
      float d = sqdist<DIMS>(sp + p * DIMS, c);
      if (s > 1) d = fminf(sd[p], d);
      sd[p] = d;
      unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)(0xffffffffu - (unsigned)p);
      best = key > best ? key : best;
    }
    best = warp_max_u64(best);
    if (lane == 0) slot[s & 1][warp] = best;
    __syncthreads();
    unsigned long long v = lane < nwarps ? slot[s & 1][lane] : 0ULL;
    v = warp_max_u64(v);
    cur = (int)(0xffffffffu - (unsigned)(v & 0xffffffffu));
    if (tid == 0) idx_out[(size_t)g * m + s] = (int64_t)g * n + cur;
  }
}

// Register-resident variant for n <= 8192: every thread keeps PPT points (coordinates and running min-distance)
// in registers, the per-sample arg-max is two redux.sync per warp (max of the distance bits, then min of the
// indices that attain it -- the same order as the 64-bit key above: largest distance, lowest index) and one
// barrier across the (few) warps.  ~0.15 us per sample instead of ~0.6 us.
template <int DIMS, int PPT, int MAXT>
__global__ void __launch_bounds__(MAXT) fps_reg_kernel(const float* __restrict__ pos, int n, int m,
                                                       int64_t* __restrict__ idx_out) {
  extern __shared__ __align__(16) float smem[];
  float* sp = smem;                 // [n][DIMS] (winner lookup)
  __shared__ unsigned slot_d[2][32], slot_i[2][32];
  const int g = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const float* gp = pos + (size_t)g * n * DIMS;
  for (int i = tid; i < n * DIMS; i += nt) sp[i] = gp[i];
  __syncthreads();
  float px[PPT][DIMS], md[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int p = tid + i * nt;
#pragma unroll
    for (int d = 0; d < DIMS; ++d) px[i][d] = p < n ? sp[p * DIMS + d] : 0.0f;
    md[i] = p < n ? INFINITY : 0.0f;          // padding slots stay at distance 0 with the highest index: never chosen
  }
  int cur = 0;
  if (tid == 0) idx_out[(size_t)g * m] = (int64_t)g * n;
  for (int s = 1; s < m; ++s) {
    float c[DIMS];
#pragma unroll
    for (int d = 0; d < DIMS; ++d) c[d] = sp[cur * DIMS + d];
    unsigned bd = 0u, bi = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const float d = fminf(md[i], sqdist<DIMS>(px[i], c));
      md[i] = d;
      const unsigned db = __float_as_uint(d);
      const unsigned p = (unsigned)(tid + i * nt);
      // points of a thread are visited in increasing index: the first maximum wins (also when every distance is 0)
      if (p < (unsigned)n && (db > bd || bi == 0xffffffffu)) { bd = db; bi = p; }
    }
    const unsigned wd = __reduce_max_sync(0xffffffffu, bd);
    const unsigned wi = __reduce_min_sync(0xffffffffu, bd == wd ? bi : 0xffffffffu);
    if (MAXT == 32) {             // a single warp holds the whole geometry: no barrier, no exchange through shared memory
      cur = (int)wi;
    } else {
      if (lane == 0) { slot_d[s & 1][warp] = wd; slot_i[s & 1][warp] = wi; }
      __syncthreads();
      unsigned vd = lane < nwarps ? slot_d[s & 1][lane] : 0u;
      unsigned vi = lane < nwarps ? slot_i[s & 1][lane] : 0xffffffffu;
      const unsigned gd = __reduce_max_sync(0xffffffffu, vd);
      cur = (int)__reduce_min_sync(0xffffffffu, vd == gd ? vi : 0xffffffffu);
    }
    if (tid == 0) idx_out[(size_t)g * m + s] = (int64_t)g * n + cur;
  }
}

// Thread-block-cluster variant for 2048 < n <= 65536: a geometry is spread over the CTAs of one cluster (contiguous
// index ranges, PPT points per thread in registers).  Per sample every CTA reduces its own range (warp redux + one block
// barrier), the thread that owns the CTA's best point publishes {distance, index, coordinates} in its CTA's shared
// memory, ONE cluster barrier later every thread reads the (<= 8) published records of all CTAs through distributed
// shared memory and picks the winner -- so the per-thread chain stays as short as in the single-CTA kernel while up to
// 8 SMs work on one geometry.  Same arithmetic and tie rule (largest distance, then lowest index).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_ld_u32(const void* local_smem_ptr, uint32_t cta_rank) {
  uint32_t laddr = static_cast<uint32_t>(__cvta_generic_to_shared(local_smem_ptr)), raddr, v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(cta_rank));
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(raddr) : "memory");
  return v;
}

template <int DIMS, int PPT>
__global__ void __launch_bounds__(1024) fps_cluster_kernel(const float* __restrict__ pos, int n, int m, int csize,
                                                           int chunk, int64_t* __restrict__ idx_out) {
  __shared__ unsigned slot_d[2][32], slot_i[2][32];
  __shared__ unsigned rec[2][2 + DIMS];                 // published record of this CTA: distance bits, index, coordinates
  const int g = blockIdx.x / csize;
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const float* gp = pos + (size_t)g * n * DIMS;
  const int p_lo = (int)rank * chunk;
  const int p_hi = min(n, p_lo + chunk);
  float px[PPT][DIMS], md[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int p = p_lo + tid + i * nt;
#pragma unroll
    for (int d = 0; d < DIMS; ++d) px[i][d] = p < p_hi ? __ldg(gp + (size_t)p * DIMS + d) : 0.0f;
    md[i] = p < p_hi ? INFINITY : 0.0f;
  }
  float c[DIMS];
#pragma unroll
  for (int d = 0; d < DIMS; ++d) c[d] = __ldg(gp + d);        // first sample: point 0
  if (rank == 0 && tid == 0) idx_out[(size_t)g * m] = (int64_t)g * n;
  for (int s = 1; s < m; ++s) {
    unsigned bd = 0u, bi = 0xffffffffu;
    int bslot = 0;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const float d = fminf(md[i], sqdist<DIMS>(px[i], c));
      md[i] = d;
      const unsigned db = __float_as_uint(d);
      const unsigned p = (unsigned)(p_lo + tid + i * nt);
      if (p < (unsigned)p_hi && (db > bd || bi == 0xffffffffu)) { bd = db; bi = p; bslot = i; }
    }
    const unsigned wd = __reduce_max_sync(0xffffffffu, bd);
    const unsigned wi = __reduce_min_sync(0xffffffffu, bd == wd ? bi : 0xffffffffu);
    if (lane == 0) { slot_d[s & 1][warp] = wd; slot_i[s & 1][warp] = wi; }
    __syncthreads();
    const unsigned vd = lane < nwarps ? slot_d[s & 1][lane] : 0u;
    const unsigned vi = lane < nwarps ? slot_i[s & 1][lane] : 0xffffffffu;
    const unsigned cd = __reduce_max_sync(0xffffffffu, vd);
    const unsigned cidx = __reduce_min_sync(0xffffffffu, vd == cd ? vi : 0xffffffffu);
    if (cidx == 0xffffffffu) {                               // a CTA whose range is empty publishes "nothing"
      if (tid == 0) { rec[s & 1][0] = 0u; rec[s & 1][1] = 0xffffffffu; }
    } else if (bi == cidx) {                                 // exactly one thread owns that point
      rec[s & 1][0] = cd;
      rec[s & 1][1] = cidx;
#pragma unroll
      for (int d = 0; d < DIMS; ++d) {
        float v = px[0][d];
#pragma unroll
        for (int i = 1; i < PPT; ++i) v = bslot == i ? px[i][d] : v;
        rec[s & 1][2 + d] = __float_as_uint(v);
      }
    }
    cluster_sync_all();
    unsigned gd = 0u, gi = 0xffffffffu;
    int grank = 0;
    for (int r = 0; r < csize; ++r) {
      const unsigned d_r = dsmem_ld_u32(&rec[s & 1][0], (uint32_t)r);
      const unsigned i_r = dsmem_ld_u32(&rec[s & 1][1], (uint32_t)r);
      if (i_r != 0xffffffffu && (d_r > gd || gi == 0xffffffffu || (d_r == gd && i_r < gi))) { gd = d_r; gi = i_r; grank = r; }
    }
#pragma unroll
    for (int d = 0; d < DIMS; ++d) c[d] = __uint_as_float(dsmem_ld_u32(&rec[s & 1][2 + d], (uint32_t)grank));
    if (rank == 0 && tid == 0) idx_out[(size_t)g * m + s] = (int64_t)g * n + (int)gi;
  }
  cluster_sync_all();       // nobody leaves while a peer may still read its shared memory
}

// Point sets too large for one SM's shared memory (config 5: up to 64k boundary points per geometry): coordinates
// are read through L1/L2 and the running min-distances live in a caller-provided scratch array [n_geom][n].
// Same arithmetic and tie rule as above; ~1 us per sample, one CTA of 1024 threads per geometry.
template <int DIMS>
__global__ void __launch_bounds__(1024) fps_global_kernel(const float* __restrict__ pos, int n, int m,
                                                          int64_t* __restrict__ idx_out, float* __restrict__ dist) {
  __shared__ unsigned slot_d[2][32], slot_i[2][32];
  const int g = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const float* gp = pos + (size_t)g * n * DIMS;
  float* sd = dist + (size_t)g * n;
  int cur = 0;
  if (tid == 0) idx_out[(size_t)g * m] = (int64_t)g * n;
  for (int s = 1; s < m; ++s) {
    float c[DIMS];
#pragma unroll
    for (int d = 0; d < DIMS; ++d) c[d] = __ldg(gp + (size_t)cur * DIMS + d);
    unsigned bd = 0u, bi = 0xffffffffu;
    for (int p = tid; p < n; p += nt) {
      float pp[DIMS];
#pragma unroll
      for (int d = 0; d < DIMS; ++d) pp[d] = __ldg(gp + (size_t)p * DIMS + d);
      float d = sqdist<DIMS>(pp, c);
      if (s > 1) d = fminf(sd[p], d);
      sd[p] = d;
      const unsigned db = __float_as_uint(d);
      if (db > bd || bi == 0xffffffffu) { bd = db; bi = (unsigned)p; }
    }
    const unsigned wd = __reduce_max_sync(0xffffffffu, bd);
    const unsigned wi = __reduce_min_sync(0xffffffffu, bd == wd ? bi : 0xffffffffu);
    if (lane == 0) { slot_d[s & 1][warp] = wd; slot_i[s & 1][warp] = wi; }
    __syncthreads();
    const unsigned vd = lane < nwarps ? slot_d[s & 1][lane] : 0u;
    const unsigned vi = lane < nwarps ? slot_i[s & 1][lane] : 0xffffffffu;
    const unsigned gd = __reduce_max_sync(0xffffffffu, vd);
    cur = (int)__reduce_min_sync(0xffffffffu, vd == gd ? vi : 0xffffffffu);
    if (tid == 0) idx_out[(size_t)g * m + s] = (int64_t)g * n + cur;
  }
}

// One warp per centroid: scan the geometry's points 32 at a time, keep the first k hits by index.
template <int DIMS>
__global__ void __launch_bounds__(256) ball_query_kernel(const float* __restrict__ pos,
                                                         const int64_t* __restrict__ cidx, int n_geom, int n,
                                                         int m, float r2, int k, int32_t* __restrict__ nbr,
                                                         int32_t* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= (int64_t)n_geom * m) return;
  const int g = (int)(q / m);
  float c[DIMS];
  const int64_t ci = cidx[q];
#pragma unroll
  for (int d = 0; d < DIMS; ++d) c[d] = __ldg(pos + ci * DIMS + d);
  const float* gp = pos + (size_t)g * n * DIMS;
  int found = 0;
  for (int base = 0; base < n && found < k; base += 32) {
    const int p = base + lane;
    bool hit = false;
    if (p < n) {
      float pp[DIMS];
#pragma unroll
      for (int d = 0; d < DIMS; ++d) pp[d] = __ldg(gp + (size_t)p * DIMS + d);
      hit = sqdist<DIMS>(pp, c) < r2;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, hit);
    const int rank = found + __popc(mask & ((1u << lane) - 1u));
    if (hit && rank < k) nbr[q * k + rank] = g * n + p;
    found += __popc(mask);
  }
  if (found > k) found = k;
  for (int j = found + lane; j < k; j += 32) nbr[q * k + j] = -1;
  if (lane == 0 && count != nullptr) count[q] = found;
}

__global__ void sa_edges_kernel(const int32_t* __restrict__ nbr, int64_t m_total, int k, int64_t n_points_total,
                                int32_t* __restrict__ slots) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m_total) return;
  const int kp = k + 1;
  int w = 0;
  for (int j = 0; j < k; ++j) {
    const int v = nbr[i * k + j];
    if (v >= 0 && (int64_t)v != i) slots[i * kp + w++] = v;   // remove_self_loops: numerically equal indices
  }
  if (i < n_points_total) slots[i * kp + w++] = (int32_t)i;    // add_self_loops(num_nodes = min(n_src, n_dst))
  for (; w < kp; ++w) slots[i * kp + w] = -1;
}

// Set-abstraction geometry of a batch from a per-geometry cache (DeviceFoamDataset.build_geometry_cache): FPS centroids
// and ball-query neighbours are stored with indices LOCAL to their geometry; this kernel rebases them to the flattened
// point numbering of the batch (geometry g of the batch owns points [g*n, (g+1)*n)) and applies the same self-loop rule as
// sa_edges_kernel (which depends on the position of a geometry inside the batch, so slots cannot be cached themselves).
__global__ void sa_cached_geometry_kernel(const int64_t* __restrict__ idx_local, const int32_t* __restrict__ nbr_local,
                                          int64_t m_total, int m, int k, int n, int64_t n_points_total,
                                          int64_t* __restrict__ idx, int32_t* __restrict__ slots) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m_total) return;
  const int64_t base = (i / m) * n;
  idx[i] = idx_local[i] + base;
  const int kp = k + 1;
  int w = 0;
  for (int j = 0; j < k; ++j) {
    const int v = nbr_local[i * k + j];
    if (v >= 0 && (int64_t)v + base != i) slots[i * kp + w++] = (int32_t)(v + base);
  }
  if (i < n_points_total) slots[i * kp + w++] = (int32_t)i;
  for (; w < kp; ++w) slots[i * kp + w] = -1;
}

// one thread per (edge, chunk of 4 output columns): 16-byte stores into the 16-byte aligned rows of ein, and a 16-byte
// load where the chunk lies inside the feature block of an aligned x row (the 128-wide level-1 features); pad columns = 0
__global__ void __launch_bounds__(256) sa_gather_kernel(const float* __restrict__ x, int ldx, int f_in, const float* __restrict__ pos,
                                 int dims, const int64_t* __restrict__ cidx, const int32_t* __restrict__ slots,
                                 int64_t n_edges, int kp, float r, float* __restrict__ ein, int ldein, int x_vec) {
  const int width = f_in + dims;
  const int chunks = ldein >> 2;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_edges * chunks) return;
  const int64_t e = t / chunks;
  const int col0 = (int)(t - e * chunks) * 4;
  const int j = __ldg(slots + e);
  float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  if (j >= 0) {
    if (x_vec && col0 + 4 <= f_in) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(x + (int64_t)j * ldx + col0));
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
      int64_t ci = -1;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int col = col0 + u;
        if (col < f_in) {
          v[u] = __ldg(x + (int64_t)j * ldx + col);
        } else if (col < width) {
          const int d = col - f_in;
          if (ci < 0) ci = __ldg(cidx + e / kp);
          // reference operator precedence: pos_j - (pos_i / r)   (models/modules.py:287)
          v[u] = __fsub_rn(__ldg(pos + (int64_t)j * dims + d), __fdiv_rn(__ldg(pos + ci * dims + d), r));
        }
      }
    }
  }
  *reinterpret_cast<float4*>(ein + e * ldein + col0) = make_float4(v[0], v[1], v[2], v[3]);
}

__global__ void sa_scatter_bwd_kernel(const float* __restrict__ gein, int ldgein, const int32_t* __restrict__ slots,
                                      int64_t n_edges, int f_in, float* gx, int ldgx) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_edges * f_in) return;
  const int64_t e = t / f_in;
  const int col = (int)(t % f_in);
  const int j = slots[e];
  if (j >= 0) atomicAdd(gx + (int64_t)j * ldgx + col, gein[e * ldgein + col]);
}

struct ColList { int32_t c[32]; };

__global__ void gather_cols_kernel(const float* __restrict__ data, int64_t n_rows, int f,
                                   const int64_t* __restrict__ row_ids, int64_t first_row, int64_t n_sel,
                                   ColList cols, int identity_cols, int n_cols, int64_t total,
                                   float* __restrict__ out, int ldout,
                                   int64_t out_rows_per_geom, int64_t out_row_offset, int out_col_offset) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % n_cols);
  const int64_t gi = t / n_cols;
  const int64_t i = gi % n_sel, g = gi / n_sel;
  const int64_t src_row = row_ids != nullptr ? row_ids[g * n_sel + i] : first_row + i;
  out[(g * out_rows_per_geom + out_row_offset + i) * ldout + out_col_offset + c] =
      __ldg(data + (g * n_rows + src_row) * f + (identity_cols ? c : cols.c[c]));
}

// the same gather for consecutive columns of 16-byte aligned rows (the row gathers of the compacted max-pool backward,
// jets copied between buffers): one thread per (row, 4 columns), 16-byte loads and stores
__global__ void __launch_bounds__(256) gather_rows_vec_kernel(const float* __restrict__ data, int64_t n_rows, int f,
                                                              const int64_t* __restrict__ row_ids, int64_t first_row, int64_t n_sel,
                                                              int chunks, int64_t total, float* __restrict__ out, int ldout,
                                                              int64_t out_rows_per_geom, int64_t out_row_offset, int out_col_offset) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % chunks) * 4;
  const int64_t gi = t / chunks;
  const int64_t i = gi % n_sel, g = gi / n_sel;
  const int64_t src_row = row_ids != nullptr ? __ldg(row_ids + g * n_sel + i) : first_row + i;
  const float4 v = __ldg(reinterpret_cast<const float4*>(data + (g * n_rows + src_row) * f + c));
  *reinterpret_cast<float4*>(out + (g * out_rows_per_geom + out_row_offset + i) * ldout + out_col_offset + c) = v;
}

__global__ void seed_jet_kernel(const float* __restrict__ data, int64_t n_rows, int f,
                                const int64_t* __restrict__ row_ids, int64_t n_sel, ColList cols, int dims, int cj,
                                int64_t total, float* __restrict__ z, int64_t ps, int ldz) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t i = t % n_sel, g = t / n_sel;
  const int64_t src_row = row_ids != nullptr ? row_ids[g * n_sel + i] : i;
  for (int d = 0; d < dims; ++d) {
    z[t * ldz + d] = __ldg(data + (g * n_rows + src_row) * f + cols.c[d]);
    for (int c = 1; c < cj; ++c) z[c * ps + t * ldz + d] = (c == 1 + d) ? 1.0f : 0.0f;
  }
}

__global__ void advance_seed_kernel(uint64_t* seed) { *seed = mix64(*seed); }

}  // namespace pcfd

using namespace pcfd;

extern "C" int pcfd_fps_bucket_supported(int32_t n_geom, int32_t n, int32_t dims);
extern "C" size_t pcfd_fps_bucket_workspace_bytes(int32_t n_geom, int32_t n, int32_t dims);
extern "C" int pcfd_fps_bucket(const float* pos, int32_t n_geom, int32_t n, int32_t dims, int32_t m, int64_t* idx_out,
                               void* workspace, size_t workspace_bytes, void* stream);

// point sets from this size on take the bucketed sampler (fps_bucket.cu); PCFD_FPS_BUCKET_MIN overrides (0 = never)
static int fps_bucket_min() {
  static int v = -1;
  if (v < 0) { const char* ev = getenv("PCFD_FPS_BUCKET_MIN"); v = ev ? atoi(ev) : 4096; }
  return v;
}
static bool fps_use_bucket(int32_t n_geom, int32_t n, int32_t dims) {
  const int mn = fps_bucket_min();
  return mn > 0 && n >= mn && pcfd_fps_bucket_supported(n_geom, n, dims);
}

extern "C" size_t pcfd_fps_workspace_bytes(int32_t n_geom, int32_t n, int32_t dims) {
  if (n_geom <= 0 || n <= 0) return 0;
  if (fps_use_bucket(n_geom, n, dims)) return pcfd_fps_bucket_workspace_bytes(n_geom, n, dims);
  // the register / cluster kernels (n <= 65536) need none; the size is reported all the same so that the fall-back
  // (PCFD_FPS_CLUSTER=0) keeps working with the caller's buffer
  return (size_t)n * (dims + 1) * sizeof(float) > 220 * 1024 ? (size_t)n_geom * n * sizeof(float) : 0;
}

extern "C" int pcfd_fps_ws(const float* pos, int32_t n_geom, int32_t n, int32_t dims, int32_t m, int64_t* idx_out,
                           void* workspace, size_t workspace_bytes, void* stream) {
  if (!pos || !idx_out || n_geom <= 0 || n <= 0 || m <= 0 || m > n || (dims != 2 && dims != 3)) return PCFD_ERR_ARG;
  // large point sets: exact bucketed sampling (needs the caller's workspace; pcfd_fps without one keeps the plain kernels)
  if (workspace != nullptr && fps_use_bucket(n_geom, n, dims))
    return pcfd_fps_bucket(pos, n_geom, n, dims, m, idx_out, workspace, workspace_bytes, stream);
  // The cluster kernel is bit-exact (tests run it with PCFD_FPS_CLUSTER=1) but NOT the default: measured on a B200, the
  // per-sample cluster barrier + distributed-shared-memory reads cost more than the shorter per-thread chain saves
  // (windbreaks 8192 -> 4096 points, 2 geometries: 8.8 ms against 5.3 ms for one CTA per geometry; 16384 points: 25 vs 18).
  static int use_cluster = -1;
  if (use_cluster < 0) { const char* ev = getenv("PCFD_FPS_CLUSTER"); use_cluster = ev ? atoi(ev) : 0; }
  if (use_cluster && n > 2048 && n <= 65536) {
    // cluster size: 2 points per thread while 8 CTAs x 1024 threads suffice, then more points per thread
    int csize = (n + 2047) / 2048;
    csize = csize <= 2 ? 2 : (csize <= 4 ? 4 : 8);
    const int chunk = (n + csize - 1) / csize;
    const int ppt = chunk <= 2048 ? 2 : (chunk <= 4096 ? 4 : 8);
    int threads = ((chunk + ppt - 1) / ppt + 31) / 32 * 32;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_geom * csize));
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t ce;
#define PCFD_FPS_CL(D_, P_) ce = cudaLaunchKernelEx(&cfg, fps_cluster_kernel<D_, P_>, pos, (int)n, (int)m, csize, chunk, idx_out)
    if (dims == 2) { if (ppt == 2) PCFD_FPS_CL(2, 2); else if (ppt == 4) PCFD_FPS_CL(2, 4); else PCFD_FPS_CL(2, 8); }
    else { if (ppt == 2) PCFD_FPS_CL(3, 2); else if (ppt == 4) PCFD_FPS_CL(3, 4); else PCFD_FPS_CL(3, 8); }
#undef PCFD_FPS_CL
    if (ce != cudaSuccess) return PCFD_ERR_CUDA + (int)ce;
    return PCFD_OK;
  }
  const size_t smem = (size_t)n * (dims + 1) * sizeof(float);
  if (smem > 220 * 1024) {
    if (workspace == nullptr || workspace_bytes < pcfd_fps_workspace_bytes(n_geom, n, dims)) return PCFD_ERR_WORKSPACE;
    if (dims == 2) fps_global_kernel<2><<<n_geom, 1024, 0, (cudaStream_t)stream>>>(pos, n, m, idx_out, (float*)workspace);
    else fps_global_kernel<3><<<n_geom, 1024, 0, (cudaStream_t)stream>>>(pos, n, m, idx_out, (float*)workspace);
    PCFD_CHECK_LAUNCH();
    return PCFD_OK;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (n <= 8192) {
    const size_t rsmem = (size_t)n * dims * sizeof(float);
    // points per thread: as few as the 1024-thread limit allows, but 2 rather than 1 (measured at n = 1000, 32 geometries:
    // 1 -> 0.152 ms, 2 -> 0.135, 4 -> 0.204, 8 -> 0.292, 16 -> 0.577, one warp with 32 -> 0.333: the per-thread update /
    // compare chain, not the block barrier, is the critical path of a sample)
    static int force_ppt = -1;
    if (force_ppt < 0) { const char* ev = getenv("PCFD_FPS_PPT"); force_ppt = ev ? atoi(ev) : 0; }
    int ppt = n <= 2048 ? 2 : (n <= 4096 ? 4 : 8);
    if ((force_ppt == 1 || force_ppt == 2 || force_ppt == 4 || force_ppt == 8 || force_ppt == 16) && n <= 1024 * force_ppt)
      ppt = force_ppt;
    int rthreads = (n + ppt - 1) / ppt;
    rthreads = (rthreads + 31) / 32 * 32;
    if (rthreads < 32) rthreads = 32;
#define PCFD_FPS_REG(D_, P_)                                                                                             \
    {                                                                                                                    \
      e = cudaFuncSetAttribute(fps_reg_kernel<D_, P_, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);   \
      if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;                                                               \
      fps_reg_kernel<D_, P_, 1024><<<n_geom, rthreads, rsmem, st>>>(pos, n, m, idx_out);                                 \
    }
    static int one_warp = -1;     // PCFD_FPS_ONE_WARP=0 disables the single-warp variant (experiments)
    if (one_warp < 0) { const char* ev = getenv("PCFD_FPS_ONE_WARP"); one_warp = ev ? atoi(ev) : 0; }   // measured slower: 0.33 vs 0.29 ms
    if (n <= 1024 && one_warp) {
      // one warp per geometry, 32 points per lane: the per-sample critical path loses the block barrier
      if (dims == 2) {
        e = cudaFuncSetAttribute(fps_reg_kernel<2, 32, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);
        if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;
        fps_reg_kernel<2, 32, 32><<<n_geom, 32, rsmem, st>>>(pos, n, m, idx_out);
      } else {
        e = cudaFuncSetAttribute(fps_reg_kernel<3, 32, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);
        if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;
        fps_reg_kernel<3, 32, 32><<<n_geom, 32, rsmem, st>>>(pos, n, m, idx_out);
      }
      PCFD_CHECK_LAUNCH();
      return PCFD_OK;
    }
    if (dims == 2) {
      if (ppt == 1) PCFD_FPS_REG(2, 1) else if (ppt == 2) PCFD_FPS_REG(2, 2) else if (ppt == 4) PCFD_FPS_REG(2, 4)
      else if (ppt == 8) PCFD_FPS_REG(2, 8) else PCFD_FPS_REG(2, 16)
    } else {
      if (ppt == 1) PCFD_FPS_REG(3, 1) else if (ppt == 2) PCFD_FPS_REG(3, 2) else if (ppt == 4) PCFD_FPS_REG(3, 4)
      else if (ppt == 8) PCFD_FPS_REG(3, 8) else PCFD_FPS_REG(3, 16)
    }
#undef PCFD_FPS_REG
    PCFD_CHECK_LAUNCH();
    return PCFD_OK;
  }
  int threads = n <= 1024 ? 256 : (n <= 4096 ? 512 : 1024);
  if (dims == 2) {
    e = cudaFuncSetAttribute(fps_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;
    fps_kernel<2><<<n_geom, threads, smem, st>>>(pos, n, m, idx_out);
  } else {
    e = cudaFuncSetAttribute(fps_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;
    fps_kernel<3><<<n_geom, threads, smem, st>>>(pos, n, m, idx_out);
  }
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

extern "C" int pcfd_fps(const float* pos, int32_t n_geom, int32_t n, int32_t dims, int32_t m, int64_t* idx_out,
                        void* stream) {
  return pcfd_fps_ws(pos, n_geom, n, dims, m, idx_out, nullptr, 0, stream);
}

extern "C" int pcfd_ball_query(const float* pos, const int64_t* centroid_idx, int32_t n_geom, int32_t n, int32_t dims,
                               int32_t m, float r, int32_t k, int32_t* nbr, int32_t* count, void* stream) {
  if (!pos || !centroid_idx || !nbr || n_geom <= 0 || n <= 0 || m <= 0 || k <= 0 || (dims != 2 && dims != 3))
    return PCFD_ERR_ARG;
  const int64_t q = (int64_t)n_geom * m;
  const unsigned blocks = (unsigned)((q + 7) / 8);
  const float r2 = r * r;
  cudaStream_t st = (cudaStream_t)stream;
  if (dims == 2) ball_query_kernel<2><<<blocks, 256, 0, st>>>(pos, centroid_idx, n_geom, n, m, r2, k, nbr, count);
  else ball_query_kernel<3><<<blocks, 256, 0, st>>>(pos, centroid_idx, n_geom, n, m, r2, k, nbr, count);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

extern "C" int pcfd_sa_edges(const int32_t* nbr, int64_t m_total, int32_t k, int64_t n_points_total, int32_t* slots,
                             void* stream) {
  if (!nbr || !slots || m_total <= 0 || k <= 0) return PCFD_ERR_ARG;
  sa_edges_kernel<<<(unsigned)((m_total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(nbr, m_total, k, n_points_total, slots);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

extern "C" int pcfd_sa_cached_geometry(const int64_t* idx_local, const int32_t* nbr_local, int32_t n_geom, int32_t m,
                                       int32_t k, int32_t n, int64_t* idx, int32_t* slots, void* stream) {
  if (!idx_local || !nbr_local || !idx || !slots || n_geom <= 0 || m <= 0 || k <= 0 || n <= 0) return PCFD_ERR_ARG;
  if ((int64_t)n_geom * n >= ((int64_t)1 << 31)) return PCFD_ERR_ARG;
  const int64_t m_total = (int64_t)n_geom * m;
  sa_cached_geometry_kernel<<<(unsigned)((m_total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      idx_local, nbr_local, m_total, m, k, n, (int64_t)n_geom * n, idx, slots);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

extern "C" int pcfd_sa_gather(const float* x, int32_t ldx, int32_t f_in, const float* pos, int32_t dims,
                              const int64_t* centroid_idx, const int32_t* slots, int64_t m_total, int32_t kp, float r,
                              float* ein, int32_t ldein, void* stream) {
  if (!pos || !centroid_idx || !slots || !ein || m_total <= 0 || kp <= 0 || f_in < 0 || (f_in > 0 && !x))
    return PCFD_ERR_ARG;
  if (ldein % 4 != 0 || ldein < f_in + dims || (reinterpret_cast<uintptr_t>(ein) & 15) != 0) return PCFD_ERR_ARG;
  const int64_t total = m_total * kp * (ldein / 4);
  const int x_vec = x != nullptr && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  sa_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x, ldx, f_in, pos, dims, centroid_idx, slots, m_total * kp, kp, r, ein, ldein, x_vec);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

extern "C" int pcfd_sa_scatter_bwd(const float* gein, int32_t ldgein, const int32_t* slots, int64_t m_total,
                                   int32_t kp, int32_t f_in, float* gx, int32_t ldgx, void* stream) {
  if (!gein || !slots || !gx || m_total <= 0 || kp <= 0 || f_in <= 0) return PCFD_ERR_ARG;
  const int64_t total = m_total * kp * f_in;
  sa_scatter_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      gein, ldgein, slots, m_total * kp, f_in, gx, ldgx);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

extern "C" int pcfd_gather_cols(const float* data, int32_t n_geom, int64_t n_rows, int32_t f, const int64_t* row_ids,
                                int64_t first_row, int64_t n_sel, const int32_t* cols_host, int32_t n_cols,
                                float* out, int32_t ldout, int64_t out_rows_per_geom, int64_t out_row_offset,
                                int32_t out_col_offset, void* stream) {
  if (!data || !out || n_cols <= 0 || n_cols > f || n_geom <= 0 || n_sel < 0) return PCFD_ERR_ARG;
  if (cols_host != nullptr && n_cols > 32) return PCFD_ERR_ARG;
  if (n_sel == 0) return PCFD_OK;
  ColList cl;
  for (int i = 0; i < 32; ++i) cl.c[i] = 0;
  if (cols_host != nullptr)
    for (int i = 0; i < n_cols; ++i) {
      if (cols_host[i] < 0 || cols_host[i] >= f) return PCFD_ERR_ARG;
      cl.c[i] = cols_host[i];
    }
  // identity columns in whole 16-byte chunks of aligned rows: the vector form
  if (cols_host == nullptr && n_cols % 4 == 0 && n_cols >= 4 && f % 4 == 0 && ldout % 4 == 0 && out_col_offset % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(data) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const int chunks = n_cols / 4;
    const int64_t total4 = (int64_t)n_geom * n_sel * chunks;
    gather_rows_vec_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        data, n_rows, f, row_ids, first_row, n_sel, chunks, total4, out, ldout, out_rows_per_geom, out_row_offset, out_col_offset);
    PCFD_CHECK_LAUNCH();
    return PCFD_OK;
  }
  const int64_t total = (int64_t)n_geom * n_sel * n_cols;
  gather_cols_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      data, n_rows, f, row_ids, first_row, n_sel, cl, cols_host == nullptr ? 1 : 0, n_cols, total, out, ldout,
      out_rows_per_geom, out_row_offset, out_col_offset);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

extern "C" int pcfd_seed_jet(const float* data, int32_t n_geom, int64_t n_rows, int32_t f, const int64_t* row_ids,
                             int64_t n_sel, const int32_t* coord_cols_host, int32_t dims, int32_t cj, float* zout,
                             int64_t plane_stride, int32_t ldz, void* stream) {
  if (!data || !zout || !coord_cols_host || (dims != 2 && dims != 3) || !valid_cj(cj) || n_sel <= 0) return PCFD_ERR_ARG;
  ColList cl;
  for (int i = 0; i < dims; ++i) cl.c[i] = coord_cols_host[i];
  const int64_t total = (int64_t)n_geom * n_sel;
  seed_jet_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      data, n_rows, f, row_ids, n_sel, cl, dims, cj, total, zout, plane_stride, ldz);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}

extern "C" int pcfd_zero(float* p, int64_t n, void* stream) {
  if (n <= 0) return PCFD_OK;
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)n * sizeof(float), (cudaStream_t)stream);
  return e == cudaSuccess ? PCFD_OK : PCFD_ERR_CUDA + (int)e;
}

extern "C" int pcfd_advance_seed(uint64_t* seed_dev, void* stream) {
  if (!seed_dev) return PCFD_ERR_ARG;
  advance_seed_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(seed_dev);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}
