// Backward of "last MLP layer -> max pool" without the dense detour.
//
// The value-only encoders of the reference end in a max over points / neighbours
//   out[s][c] = max_j act(z[s*L + j][c]),   z = T(zin) W^T + b
// (set-abstraction convolution models/modules.py:286-292 + 313-325, GlobalSetAbstraction :412-423, Branch :184-190,
// GeometryEncoder :206-214).  The cotangent of z therefore has ONE non-zero entry per (segment, channel): row
// arg[s][c].  autograd materialises it as a dense (rows x C) tensor and runs two dense GEMMs over it; here
//
//   pool_dw_kernel   gw[c][:] += g[s][c] * T(zin)[s*L + arg[s][c]][:],  gbias[c] += g[s][c]
//                    (C*k multiply-adds per segment instead of L*C*k), per-CTA partial sums in registers over a fixed
//                    set of segments, reduced in a fixed order by the common dW finish kernel (deterministic).  The
//                    activated input rows of a BATCH of segments are staged in shared memory at once, so that the
//                    DRAM latency of a batch is paid once and not once per segment.
//   pool_dx_kernel   gzin[s*L + j][:] = T'(zin) * sum_{c: arg[s][c] == j} g[s][c] * W[c][:]
//                    one warp per row; the channels that selected the row are enumerated in channel order with ballots
//                    (deterministic), W rows stream through L1; the row sums go through a shared tile so that the
//                    final pass (activation derivative, 16-byte stores) has all its zin loads in flight together
//   pool_compact_kernel   for long segments (L > C): only <= C rows of a segment carry gradient, so the remaining
//                    layers of the encoder run their ordinary dense backward on those rows alone: emits the row ids and
//                    the (diagonal) cotangent of the compacted rows
//
// with g[s][c] = gout[s][c] * act'(zsel[s][c]), zsel = the pre-activation of the selected row (written by the forward
// max, pcfd_segmax_fwd_z).  CUDA-core kernels: the work is 1/L of the dense form and bound by HBM / shared-memory
// bandwidth, not by the tensor pipe.
#include "common.cuh"

namespace pcfd {

constexpr int PB_THREADS = 256;
constexpr int PB_TCH = 16;           // channels per thread (register accumulators: PB_TCH x 4 columns)
constexpr int PB_TILE_BYTES = 36 * 1024;   // shared-memory budget of a batch of staged segments (when one segment fits)

__device__ __forceinline__ float4 act4(int act, float4 v) {
  if (act != PCFD_ACT_NONE) { v.x = act_value(act, v.x); v.y = act_value(act, v.y); v.z = act_value(act, v.z); v.w = act_value(act, v.w); }
  return v;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// grid (splits, channel blocks).  Dynamic shared memory: `nstages` stages of {tile [gb][seg_len][kt4] float4, raw arg /
// gout / zsel [gb][cb]} filled by cp.async one batch ahead of the arithmetic, and gr [gb][cbp] (cotangent, tile row).
template <int NIT>
__global__ void __launch_bounds__(PB_THREADS) pool_dw_kernel(const float* __restrict__ gout, int ldgout,
                                                             const int32_t* __restrict__ arg, const float* __restrict__ zsel,
                                                             int ldzsel, int act_pool, int64_t n_seg, int seg_len, int c,
                                                             const float* __restrict__ zin, int ldzin, int act_in, int k,
                                                             int cb, int gb, int nstages, float* __restrict__ partial,
                                                             float* __restrict__ colsum) {
  extern __shared__ __align__(16) float pb_smem[];
  const int kt4 = (k + 3) >> 2;
  const int seg_items = seg_len * kt4;
  const int t = threadIdx.x;
  const int slots = PB_THREADS / kt4;
  const int cs = t / kt4, kq = t - cs * kt4;
  const bool active = cs < slots;
  const int ch0 = blockIdx.y * cb;
  const int cbn = min(cb, c - ch0);
  const int cbp = NIT * slots;                    // channel slots per segment in `gr` (zero cotangent beyond cbn)
  const size_t stage_floats = (size_t)gb * seg_items * 4 + (size_t)gb * cb * 3;
  const size_t stage_stride = (stage_floats + 3) & ~(size_t)3;
  float2* gr = reinterpret_cast<float2*>(pb_smem + stage_stride * nstages);
  const int splits = gridDim.x;
  // this CTA's segments: a contiguous range (the rows of a batch are then one contiguous block of zin)
  const int64_t per = (n_seg + splits - 1) / splits;
  const int64_t sbeg = (int64_t)blockIdx.x * per, send = min(n_seg, sbeg + per);
  const int nb = sbeg < send ? (int)((send - sbeg + gb - 1) / gb) : 0;

  float4 acc[NIT];
  float bsum[NIT];
#pragma unroll
  for (int i = 0; i < NIT; ++i) { acc[i] = make_float4(0.f, 0.f, 0.f, 0.f); bsum[i] = 0.f; }

  auto issue = [&](int b) {
    const int64_t s0 = sbeg + (int64_t)b * gb;
    const int ns = (int)min((int64_t)gb, send - s0);
    float* st = pb_smem + stage_stride * (b % nstages);
    const float* zrow = zin + s0 * (int64_t)seg_len * ldzin;
    const int items = ns * seg_items;
    if (ldzin == kt4 * 4) {
      for (int item = t; item < items; item += PB_THREADS) cp_async16(st + (size_t)item * 4, zrow + (size_t)item * 4);
    } else {
      for (int item = t; item < items; item += PB_THREADS) {
        const int r = item / kt4, q = item - r * kt4;
        cp_async16(st + (size_t)item * 4, zrow + (int64_t)r * ldzin + 4 * q);
      }
    }
    float* raw = st + (size_t)gb * seg_items * 4;
    for (int si = 0; si < ns; ++si) {
      const int64_t s = s0 + si;
      const int32_t* pa = arg + s * c + ch0;
      const float* pg = gout + s * ldgout + ch0;
      const float* pz = zsel + s * ldzsel + ch0;
      float* ra = raw + si * cb;
      for (int xc = t; xc < cbn; xc += PB_THREADS) {
        cp_async4(ra + xc, pa + xc);
        cp_async4(ra + gb * cb + xc, pg + xc);
        cp_async4(ra + 2 * gb * cb + xc, pz + xc);
      }
    }
    cp_async_commit();
  };

  if (nstages == 2 && nb > 0) issue(0);
  for (int b = 0; b < nb; ++b) {
    if (nstages == 2) {
      if (b + 1 < nb) { issue(b + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    } else {
      issue(b);
      cp_async_wait<0>();
    }
    __syncthreads();
    const int64_t s0 = sbeg + (int64_t)b * gb;
    const int ns = (int)min((int64_t)gb, send - s0);
    float* st = pb_smem + stage_stride * (b % nstages);
    const float* raw = st + (size_t)gb * seg_items * 4;
    // ---- activate the staged rows in place; columns beyond k are padding of the row (uninitialised memory)
    const int items = ns * seg_items;
    const bool ragged = (k & 3) != 0;
    if (act_in != PCFD_ACT_NONE || ragged) {
      for (int item = t; item < items; item += PB_THREADS) {
        float4 a = act4(act_in, *reinterpret_cast<const float4*>(st + (size_t)item * 4));
        if (ragged) {
          const int col = 4 * (item % kt4);
          if (col + 1 >= k) a.y = 0.f;
          if (col + 2 >= k) a.z = 0.f;
          if (col + 3 >= k) a.w = 0.f;
        }
        *reinterpret_cast<float4*>(st + (size_t)item * 4) = a;
      }
    }
    for (int x = t; x < ns * cbp; x += PB_THREADS) {
      const int si = x / cbp, xc = x - si * cbp;
      float g = 0.f;
      int a = 0;
      if (xc < cbn) {
        a = __float_as_int(raw[si * cb + xc]);
        g = a >= 0 ? raw[gb * cb + si * cb + xc] * act_d1(act_pool, raw[2 * gb * cb + si * cb + xc]) : 0.f;
        a = a >= 0 ? a : 0;
      }
      gr[x] = make_float2(g, __int_as_float((a + si * seg_len) * kt4 * 16));     // byte offset of the tile row
    }
    __syncthreads();
    if (active) {
      const char* tileb = reinterpret_cast<const char*>(st) + kq * 16;
      for (int si = 0; si < ns; ++si) {
        const float2* grs = gr + si * cbp + cs;
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
          const float2 ga = grs[i * slots];
          const float4 a4 = *reinterpret_cast<const float4*>(tileb + __float_as_int(ga.y));
          acc[i].x += ga.x * a4.x; acc[i].y += ga.x * a4.y; acc[i].z += ga.x * a4.z; acc[i].w += ga.x * a4.w;
          bsum[i] += ga.x;
        }
      }
    }
    __syncthreads();
  }
  if (!active) return;
#pragma unroll
  for (int i = 0; i < NIT; ++i) {
    const int chl = cs + i * slots;
    if (chl >= cbn) continue;
    const int ch = ch0 + chl;
    float* o = partial + ((size_t)blockIdx.x * c + ch) * k + 4 * kq;
    const int col = 4 * kq;
    if (col < k) o[0] = acc[i].x;
    if (col + 1 < k) o[1] = acc[i].y;
    if (col + 2 < k) o[2] = acc[i].z;
    if (col + 3 < k) o[3] = acc[i].w;
    if (kq == 0 && colsum != nullptr) colsum[(size_t)blockIdx.x * c + ch] = bsum[i];
  }
}

constexpr int PB_DX_ROWS = 64;   // rows per CTA

// A group of LG lanes (LG = 2..32, kt4 <= LG * VPL) owns a row: it scans the segment's arg-max table LG channels at a
// time (one full-warp ballot serves all groups of the warp), adds g * W[ch][:] for the channels that selected its row
// (16-byte W loads through L1 when VEC), and leaves the row sums in a shared tile.  Dynamic shared memory: per spanned
// segment c x (g, arg), then the row-sum tile [64][kt4] float4.
template <int VPL, bool VEC>
__global__ void __launch_bounds__(PB_THREADS) pool_dx_kernel(const float* __restrict__ gout, int ldgout,
                                                             const int32_t* __restrict__ arg, const float* __restrict__ zsel,
                                                             int ldzsel, int act_pool, int64_t n_seg, int seg_len, int c,
                                                             const float* __restrict__ zin, int ldzin, int act_in, int k,
                                                             const float* __restrict__ w, int ldw, float* __restrict__ gzin,
                                                             int ldgzin, int nspan_max, int lg_shift) {
  extern __shared__ __align__(16) float pb_smem[];
  const int kt4 = (k + 3) >> 2;
  const int64_t total = n_seg * seg_len;
  const int64_t r0 = (int64_t)blockIdx.x * PB_DX_ROWS;
  const int nrows = (int)min((int64_t)PB_DX_ROWS, total - r0);
  const int64_t s0 = r0 / seg_len;
  const int j0 = (int)(r0 - s0 * seg_len);
  const int nspan = (j0 + nrows - 1) / seg_len + 1;
  float* g_s = pb_smem;
  int* a_s = reinterpret_cast<int*>(pb_smem + (size_t)nspan_max * c);
  float4* acc_s = reinterpret_cast<float4*>(pb_smem + (((size_t)nspan_max * c * 2 + 3) & ~(size_t)3));
  const int nsc = nspan * c;
  for (int base = threadIdx.x; base < nsc; base += 4 * PB_THREADS) {
    int a[4];
    float go[4], zz[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x = base + u * PB_THREADS;
      if (x < nsc) {
        const int sp = x / c, ch = x - sp * c;
        const int64_t s = s0 + sp;
        a[u] = __ldg(arg + s * c + ch);
        go[u] = __ldg(gout + s * ldgout + ch);
        zz[u] = __ldg(zsel + s * ldzsel + ch);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x = base + u * PB_THREADS;
      if (x < nsc) {
        g_s[x] = a[u] >= 0 ? go[u] * act_d1(act_pool, zz[u]) : 0.f;
        a_s[x] = a[u];
      }
    }
  }
  __syncthreads();
  const int lg = 1 << lg_shift;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (lg - 1);                 // lane inside the group
  const int gshift = lane & ~(lg - 1);            // first lane of the group
  const unsigned gmask = lg == 32 ? 0xffffffffu : ((1u << lg) - 1u);
  const int group = threadIdx.x >> lg_shift;      // group index inside the CTA
  const int ngroups = PB_THREADS >> lg_shift;
  const int passes = (nrows + ngroups - 1) / ngroups;     // uniform: every lane takes part in every ballot
  for (int ps = 0; ps < passes; ++ps) {
    const int rl = ps * ngroups + group;          // row inside the CTA
    const bool live = rl < nrows;
    const int jj = j0 + rl;
    const int sp = live ? jj / seg_len : 0;
    const int j = live ? jj - sp * seg_len : -5;  // never matches
    const int* as = a_s + sp * c;
    const float* gs = g_s + sp * c;
    float4 acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int cfull = c & ~(lg - 1);
    for (int cbase = 0; cbase < c; cbase += lg) {
      const int ch = cbase + gl;
      const unsigned bal = __ballot_sync(0xffffffffu, (cbase < cfull || ch < c) && as[ch] == j);
      unsigned m = (bal >> gshift) & gmask;
      while (m) {
        const int hit = cbase + __ffs(m) - 1;
        m &= m - 1;
        const float g = gs[hit];
        const float* wr = w + (unsigned)(hit * ldw);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int q = gl + (v << lg_shift);
          if (q < kt4) {
            float4 w4;
            if (VEC) {
              w4 = __ldg(reinterpret_cast<const float4*>(w) + (unsigned)(hit * (ldw >> 2) + q));
            } else {
              const int col = 4 * q;
              w4.x = __ldg(wr + col);
              w4.y = col + 1 < k ? __ldg(wr + col + 1) : 0.f;
              w4.z = col + 2 < k ? __ldg(wr + col + 2) : 0.f;
              w4.w = col + 3 < k ? __ldg(wr + col + 3) : 0.f;
            }
            acc[v].x += g * w4.x; acc[v].y += g * w4.y; acc[v].z += g * w4.z; acc[v].w += g * w4.w;
          }
        }
      }
    }
    if (live) {
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int q = gl + (v << lg_shift);
        if (q < kt4) acc_s[(size_t)rl * kt4 + q] = acc[v];
      }
    }
  }
  __syncthreads();
  // ---- gzin = row sums * act'(zin): 16-byte loads / stores, every load of the CTA in flight at once
  const int items = nrows * kt4;
  for (int base = threadIdx.x; base < items; base += 4 * PB_THREADS) {
    float4 zv[4];
    if (act_in != PCFD_ACT_NONE) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int item = base + u * PB_THREADS;
        if (item < items) {
          const int r = item / kt4, q = item - r * kt4;
          zv[u] = __ldg(reinterpret_cast<const float4*>(zin + (r0 + r) * ldzin + 4 * q));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int item = base + u * PB_THREADS;
      if (item < items) {
        const int r = item / kt4, q = item - r * kt4;
        float4 a = acc_s[item];
        if (act_in != PCFD_ACT_NONE) {
          a.x *= act_d1(act_in, zv[u].x); a.y *= act_d1(act_in, zv[u].y); a.z *= act_d1(act_in, zv[u].z); a.w *= act_d1(act_in, zv[u].w);
        }
        const int col = 4 * q;
        if (col + 1 >= k) a.y = 0.f;
        if (col + 2 >= k) a.z = 0.f;
        if (col + 3 >= k) a.w = 0.f;
        *reinterpret_cast<float4*>(gzin + (r0 + r) * ldgzin + col) = a;
      }
    }
  }
}

// ids[s][ch] = max(arg, 0) (int64, row inside the segment); gzc[(s*c + ch)][:] = e_ch * g[s][ch]
__global__ void __launch_bounds__(PB_THREADS) pool_compact_kernel(const float* __restrict__ gout, int ldgout,
                                                                  const int32_t* __restrict__ arg,
                                                                  const float* __restrict__ zsel, int ldzsel, int act_pool,
                                                                  int64_t n_seg, int c, int64_t* __restrict__ ids,
                                                                  float* __restrict__ gzc, int ldgzc) {
  const int64_t crow = blockIdx.x;            // compact row = s*c + ch
  const int64_t s = crow / c;
  const int ch = (int)(crow - s * c);
  const int a = __ldg(arg + crow);
  float g = 0.f;
  if (a >= 0) g = __ldg(gout + s * ldgout + ch) * act_d1(act_pool, __ldg(zsel + s * ldzsel + ch));
  if (threadIdx.x == 0) ids[crow] = a >= 0 ? a : 0;
  for (int col = threadIdx.x; col < ldgzc; col += PB_THREADS) gzc[crow * ldgzc + col] = col == ch ? g : 0.f;
}

struct PoolPlan { int kt4, slots, cb, cblocks, splits, gb, nstages, nit, nspan, lg_shift, vpl; size_t smem_dw, smem_dx; bool ok; };

static PoolPlan pool_plan(int64_t n_seg, int seg_len, int k, int c) {
  PoolPlan p{};
  p.kt4 = (k + 3) / 4;
  p.ok = p.kt4 >= 1 && p.kt4 <= PB_THREADS;
  if (!p.ok) return p;
  p.slots = PB_THREADS / p.kt4;
  p.cb = p.slots * PB_TCH < c ? p.slots * PB_TCH : c;
  p.cblocks = (c + p.cb - 1) / p.cb;
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t seg_bytes = (size_t)seg_len * p.kt4 * 16;
  int gb = (int)(PB_TILE_BYTES / seg_bytes);
  if (gb < 1) gb = 1;
  if (gb > 16) gb = 16;
  // CTAs: at least two per SM when the segments allow it, but never fewer than one batch of segments each
  int64_t sp = (3 * (int64_t)sms + p.cblocks - 1) / p.cblocks;
  if (sp > n_seg) sp = n_seg;
  if (sp < 1) sp = 1;
  const int64_t per = (n_seg + sp - 1) / sp;
  if (gb > per) gb = (int)per;
  p.gb = gb;
  p.splits = (int)((n_seg + per - 1) / per);
  int nit = 1;
  while (nit * p.slots < p.cb) nit *= 2;
  p.nit = nit;
  const int cbp = nit * p.slots;
  const size_t stage = ((((size_t)gb * seg_bytes / 4 + (size_t)gb * p.cb * 3) + 3) & ~(size_t)3) * 4;
  p.nstages = (2 * stage + (size_t)gb * cbp * 8 <= 100 * 1024 && per > gb) ? 2 : 1;   // two CTAs per SM keep their double buffers
  p.smem_dw = stage * p.nstages + (size_t)gb * cbp * 8;
  p.nspan = (PB_DX_ROWS + seg_len - 2) / seg_len + 1;
  p.smem_dx = (((size_t)p.nspan * c * 2 + 3) & ~(size_t)3) * 4 + (size_t)PB_DX_ROWS * p.kt4 * 16;
  int lgs = 1;
  while ((1 << lgs) < p.kt4 && lgs < 5) ++lgs;
  p.lg_shift = lgs;
  p.vpl = (p.kt4 + (1 << lgs) - 1) >> lgs;
  p.ok = p.smem_dw <= 200 * 1024 && p.smem_dx <= 160 * 1024 && p.vpl <= 4;
  return p;
}

}  // namespace pcfd

using namespace pcfd;

extern "C" int pcfd_dw_finish(const float*, int, const float*, int32_t, float*, int32_t, float*, float*, int32_t, int64_t, int64_t,
                              int32_t, int32_t, float*, const float*, int, void*);

static bool pool_tin_ok(const pcfd_intrans_t* tin, int k) {
  if (tin == nullptr) return true;
  if (tin->escale != nullptr || tin->drop_p > 0.f) return false;
  return tin->act == PCFD_ACT_NONE || tin->act_cols <= 0 || tin->act_cols >= k;
}

extern "C" int pcfd_pool_layer_bwd_supported(int64_t n_seg, int32_t seg_len, int32_t k, int32_t c, const pcfd_intrans_t* tin_host,
                                             int32_t ldzin) {
  if (n_seg <= 0 || seg_len <= 0 || k <= 0 || c <= 0 || ldzin % 4 != 0 || !pool_tin_ok(tin_host, k)) return 0;
  return pool_plan(n_seg, seg_len, k, c).ok ? 1 : 0;
}

extern "C" size_t pcfd_pool_layer_bwd_workspace_bytes(int64_t n_seg, int32_t seg_len, int32_t k, int32_t c) {
  const PoolPlan p = pool_plan(n_seg, seg_len, k, c);
  if (!p.ok) return 0;
  return ((size_t)p.splits * c * k + (size_t)p.splits * c) * sizeof(float) + 256;
}

template <int VPL, bool VEC>
static cudaError_t launch_pool_dx(dim3 grid, size_t smem, cudaStream_t st, const float* gout, int ldgout, const int32_t* arg,
                                  const float* zsel, int ldzsel, int act_pool, int64_t n_seg, int seg_len, int c,
                                  const float* zin, int ldzin, int act_in, int k, const float* w, int ldw, float* gzin,
                                  int ldgzin, int nspan, int lg_shift) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(pool_dx_kernel<VPL, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  pool_dx_kernel<VPL, VEC><<<grid, PB_THREADS, smem, st>>>(gout, ldgout, arg, zsel, ldzsel, act_pool, n_seg, seg_len, c, zin,
                                                           ldzin, act_in, k, w, ldw, gzin, ldgzin, nspan, lg_shift);
  return cudaGetLastError();
}

extern "C" int pcfd_pool_layer_bwd(const float* gout, int32_t ldgout, const int32_t* arg, const float* zsel, int32_t ldzsel,
                                   int32_t act_pool, int64_t n_seg, int32_t seg_len, int32_t c, const float* zin,
                                   int32_t ldzin, const pcfd_intrans_t* tin_host, int32_t k, const float* w, int32_t ldw,
                                   float* gw, int32_t ldgw, float* gbias, float* gzin, int32_t ldgzin, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (!gout || !arg || !zsel || !zin || !w || n_seg <= 0 || seg_len <= 0 || c <= 0 || k <= 0) return PCFD_ERR_ARG;
  if (ldzin < k || ldw < k || ldzsel < c || ldgout < c || (gw && ldgw < k) || (gzin && (ldgzin < ((k + 3) & ~3) || ldgzin % 4)))
    return PCFD_ERR_ARG;
  if (!pcfd_pool_layer_bwd_supported(n_seg, seg_len, k, c, tin_host, ldzin)) return PCFD_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(zin) & 15) != 0 || (reinterpret_cast<uintptr_t>(gzin) & 15) != 0) return PCFD_ERR_ALIGN;
  const int rc = check_sm100();
  if (rc) return rc;
  const PoolPlan p = pool_plan(n_seg, seg_len, k, c);
  const int act_in = tin_host ? tin_host->act : PCFD_ACT_NONE;
  cudaStream_t st = (cudaStream_t)stream;
  if (gw != nullptr || gbias != nullptr) {
    if (!workspace || workspace_bytes < pcfd_pool_layer_bwd_workspace_bytes(n_seg, seg_len, k, c)) return PCFD_ERR_WORKSPACE;
    float* partial = reinterpret_cast<float*>(workspace);
    float* colsum = partial + (size_t)p.splits * c * k;
#define PCFD_POOL_DW(NIT_)                                                                                                  \
  {                                                                                                                        \
    if (p.smem_dw > 48 * 1024) {                                                                                           \
      cudaError_t e = cudaFuncSetAttribute(pool_dw_kernel<NIT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_dw); \
      if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;                                                                 \
    }                                                                                                                      \
    pool_dw_kernel<NIT_><<<dim3((unsigned)p.splits, (unsigned)p.cblocks), PB_THREADS, p.smem_dw, st>>>(                     \
        gout, ldgout, arg, zsel, ldzsel, act_pool, n_seg, seg_len, c, zin, ldzin, act_in, k, p.cb, p.gb, p.nstages, partial, \
        colsum);                                                                                                           \
  }
    switch (p.nit) {
      case 1: PCFD_POOL_DW(1) break;
      case 2: PCFD_POOL_DW(2) break;
      case 4: PCFD_POOL_DW(4) break;
      case 8: PCFD_POOL_DW(8) break;
      default: PCFD_POOL_DW(16) break;
    }
#undef PCFD_POOL_DW
    PCFD_CHECK_LAUNCH();
    const int frc = pcfd_dw_finish(gw ? partial : nullptr, p.splits, nullptr, 0, gw, ldgw, gbias, nullptr, 0, 0, 0, k, c, nullptr,
                                   colsum, p.splits, st);
    if (frc) return frc;
  }
  if (gzin != nullptr) {
    const int64_t total = n_seg * seg_len;
    const dim3 grid((unsigned)((total + PB_DX_ROWS - 1) / PB_DX_ROWS));
    cudaError_t e;
    const bool vec = ldw % 4 == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0;
#define PCFD_POOL_DX(VPL_, VEC_)                                                                                              \
  e = launch_pool_dx<VPL_, VEC_>(grid, p.smem_dx, st, gout, ldgout, arg, zsel, ldzsel, act_pool, n_seg, seg_len, c, zin, ldzin, \
                                 act_in, k, w, ldw, gzin, ldgzin, p.nspan, p.lg_shift)
    if (p.vpl <= 1) { if (vec) PCFD_POOL_DX(1, true); else PCFD_POOL_DX(1, false); }
    else if (p.vpl <= 2) { if (vec) PCFD_POOL_DX(2, true); else PCFD_POOL_DX(2, false); }
    else { if (vec) PCFD_POOL_DX(4, true); else PCFD_POOL_DX(4, false); }
#undef PCFD_POOL_DX
    if (e != cudaSuccess) return PCFD_ERR_CUDA + (int)e;
  }
  return PCFD_OK;
}

extern "C" int pcfd_pool_compact(const float* gout, int32_t ldgout, const int32_t* arg, const float* zsel, int32_t ldzsel,
                                 int32_t act_pool, int64_t n_seg, int32_t c, int64_t* ids, float* gzc, int32_t ldgzc,
                                 void* stream) {
  if (!gout || !arg || !zsel || !ids || !gzc || n_seg <= 0 || c <= 0 || ldgzc < c || ldzsel < c || ldgout < c) return PCFD_ERR_ARG;
  if (n_seg * (int64_t)c >= ((int64_t)1 << 31)) return PCFD_ERR_ARG;
  pool_compact_kernel<<<(unsigned)(n_seg * c), PB_THREADS, 0, (cudaStream_t)stream>>>(gout, ldgout, arg, zsel, ldzsel, act_pool,
                                                                                      n_seg, c, ids, gzc, ldgzc);
  PCFD_CHECK_LAUNCH();
  return PCFD_OK;
}
