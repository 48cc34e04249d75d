/*
 * pcfd.h -- C ABI of libpcfd_sm100.so: the B200 (sm_100a) kernels behind the physics-informed
 * training step of Gallinator/porous-cfd.
 *
 * The reference is pure Python and has no FFI of its own; the seams this
 * library is bound behind are the reference's Python call sites, cited per function below
 * (paths relative to the reference tree).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every function returns 0 on success or a PCFD_ERR_* code; nothing throws, nothing exits;
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in `_host`;
 *     the library never allocates, frees or retains memory; scratch space is passed in as
 *     `workspace` / `workspace_bytes` (sizes from the *_workspace_bytes queries);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no internal syncs, so
 *     every entry point is CUDA-graph capturable;
 *   - fp32 data, row-major, explicit leading dimensions (`ld*`, in elements);
 *   - "jet" tensors are stacks of `cj` channel planes [cj][rows][width]: plane 0 is the value,
 *     planes 1..D the first spatial tangents d/dx_k, planes D+1..2D the second tangents
 *     d2/dx_k2 (cj = 1, 1+D or 1+2D; D = 2 or 3).  `plane_stride` is the distance between
 *     planes in elements.
 *   - there is no CPU fallback: a device that is not sm_100 yields PCFD_ERR_ARCH.
 */
#ifndef PCFD_H_
#define PCFD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCFD_ABI_VERSION 1

enum {
  PCFD_OK = 0,
  PCFD_ERR_ARG = 1,       /* bad shape / null pointer / unsupported channel count */
  PCFD_ERR_ALIGN = 2,     /* misaligned pointer or leading dimension */
  PCFD_ERR_WORKSPACE = 3, /* workspace too small */
  PCFD_ERR_ARCH = 4,      /* current device is not sm_100 */
  PCFD_ERR_CUDA = 100     /* PCFD_ERR_CUDA + cudaError_t of a failed launch */
};

enum { PCFD_ACT_NONE = 0, PCFD_ACT_SILU = 1, PCFD_ACT_TANH = 2 };

/* ABI version, and the compute capability (major*10+minor) of the current device. */
int pcfd_abi_version(void);
int pcfd_device_arch(int* cc_out_host);
/* The library holds NO mutable state and is re-entrant: every entry point is a function of its arguments, work goes to
 * the stream passed in, the device is the caller's current device.  (Per-device records of which kernels already have
 * their shared-memory attribute set are idempotent caches behind atomics; environment switches named in DESIGN.md are
 * read once.) */

/* ------------------------------------------------------------------------------------------
 * Input transform of a jet layer.  The reference applies activation / dropout / branch scaling
 * at the END of a layer (models/modules.py:48-52 MLP; :239-245 NeuralOperator); here the layer
 * stores its pre-activations and the NEXT layer applies  a = dropout(act(z)) * escale  while it
 * loads its input, carrying the jet through it:
 *     a0 = s*f(z0)   ak = s*f'(z0)*zk   akk = s*(f''(z0)*zk^2 + f'(z0)*zkk),  s = mask/(1-p) * escale
 * ------------------------------------------------------------------------------------------ */
typedef struct pcfd_intrans {
  int32_t act;              /* PCFD_ACT_* */
  int32_t act_cols;         /* apply act/dropout/escale to the first act_cols input columns only; <=0: all */
  const float* escale;      /* [n_geom][k] per-geometry embedding (PI-GANO branch output) or NULL */
  int32_t ldescale;
  float drop_p;             /* dropout probability, 0 = off (models/modules.py:51-52, 236-237) */
  const uint64_t* seed_dev; /* device pointer to the step seed (graph-replayable); may be NULL if drop_p == 0 */
  uint32_t salt;            /* distinguishes layers that share the seed */
  uint32_t reserved;
} pcfd_intrans_t;

/*
 * Jet linear layer, forward.  Replaces torch.nn.Linear on the differentiated path
 * (models/modules.py:48, :232; models/pi_gano/pi_gano.py:47) for value AND tangent channels:
 *     zout[c] = T(zin)[c] * W^T   (+ bias + cvec[geom] on channel 0)
 * w is [n][k] (torch Linear layout) with row stride ldw, so a column block of a larger weight
 * (the per-point half of a concat layer, models/pipn/pipn_foam.py:96-98) is addressed in place.
 * cvec [n_geom][n] carries the per-geometry constant half of such a layer; geometry of row r is
 * r / rows_per_geom.
 */
int pcfd_jet_linear_fwd(const float* zin, int64_t zin_plane_stride, int32_t ldzin,
                        const pcfd_intrans_t* tin_host,
                        const float* w, int32_t ldw, const float* bias,
                        const float* cvec, int32_t ldcvec,
                        float* zout, int64_t zout_plane_stride, int32_t ldzout,
                        int32_t cj, int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n,
                        void* stream);

/*
 * Jet linear layer, backward to the layer input: reverse of the forward above including the
 * reverse of the input transform (needs f''' for second-order jets).  gzout is the gradient wrt
 * zout; gzin receives the gradient wrt zin (overwritten).  gescale (optional, [n_geom][k]) is
 * ACCUMULATED with the gradient wrt escale.  Replaces the autograd double-backward through
 * Linear/activation (models/model_base.py:11-20 create_graph=True sweeps + loss.backward()).
 */
int pcfd_jet_linear_bwd_dx(const float* gzout, int64_t gzout_plane_stride, int32_t ldgzout,
                           const float* w, int32_t ldw,
                           const float* zin, int64_t zin_plane_stride, int32_t ldzin,
                           const pcfd_intrans_t* tin_host,
                           float* gzin, int64_t gzin_plane_stride, int32_t ldgzin,
                           float* gescale, int32_t ldgescale,
                           int32_t cj, int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n,
                           void* stream);

/*
 * Jet linear layer, backward to the parameters.  All outputs are ACCUMULATED (+=):
 *     gw[n][k]   += sum_{c,rows} gzout[c][row][n] * T(zin)[c][row][k]
 *     gbias[n]   += sum_rows gzout[0][row][n]                       (optional)
 *     gcvec[g][n]+= sum_{rows of geometry g} gzout[0][row][n]        (optional)
 * The row reduction is split over CTAs and reduced in a fixed order (deterministic).
 */
size_t pcfd_jet_linear_bwd_dw_workspace_bytes(int32_t cj, int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n);
int pcfd_jet_linear_bwd_dw(const float* gzout, int64_t gzout_plane_stride, int32_t ldgzout,
                           const float* zin, int64_t zin_plane_stride, int32_t ldzin,
                           const pcfd_intrans_t* tin_host,
                           float* gw, int32_t ldgw, float* gbias, float* gcvec, int32_t ldgcvec,
                           int32_t cj, int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n,
                           void* workspace, size_t workspace_bytes, void* stream);

/*
 * Which kernel family the three entry points above execute for a given call -- a pure query, nothing is launched:
 * PCFD_ENGINE_TCGEN05 (warp-specialised TMA + tcgen05 3xTF32 kernels, the product path for every wide layer),
 * PCFD_ENGINE_THIN (fp32 streaming kernels for first / last layers and per-geometry rows, where no tensor-core tile
 * fits) or PCFD_ENGINE_FFMA (the generic fp32 CUDA-core engine: misaligned operands, fewer than 256 rows, ...).
 * pass: 0 forward (a = zin, b = zout), 1 dX (a = gzout, b = zin), 2 dW (a = gzout, b = zin).  Callers that expect their
 * wide layers on the tensor cores assert on the answer (bench.py, tests) instead of finding a slow fallback later.
 */
enum { PCFD_ENGINE_FFMA = 0, PCFD_ENGINE_THIN = 1, PCFD_ENGINE_TCGEN05 = 2 };
int pcfd_jet_linear_engine(int32_t pass, const float* a, int64_t a_plane_stride, int32_t lda, const float* w, int32_t ldw,
                           const float* b, int64_t b_plane_stride, int32_t ldb, const pcfd_intrans_t* tin_host,
                           int32_t has_gescale, int32_t cj, int64_t rows, int32_t k, int32_t n);
/*
 * The generic fp32 CUDA-core engine under its own names, same arguments as pcfd_jet_linear_{fwd,bwd_dx,bwd_dw}: an
 * independent implementation of the same layer that tests and scripts/bench_layers.py compare the tensor-core kernels
 * with (any shape, any alignment).
 */
int pcfd_ffma_jet_linear_fwd(const float* zin, int64_t zin_plane_stride, int32_t ldzin, const pcfd_intrans_t* tin_host,
                             const float* w, int32_t ldw, const float* bias, const float* cvec, int32_t ldcvec,
                             float* zout, int64_t zout_plane_stride, int32_t ldzout,
                             int32_t cj, int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n, void* stream);
int pcfd_ffma_jet_linear_bwd_dx(const float* gzout, int64_t gzout_plane_stride, int32_t ldgzout, const float* w, int32_t ldw,
                                const float* zin, int64_t zin_plane_stride, int32_t ldzin, const pcfd_intrans_t* tin_host,
                                float* gzin, int64_t gzin_plane_stride, int32_t ldgzin, float* gescale, int32_t ldgescale,
                                int32_t cj, int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n, void* stream);
int pcfd_ffma_jet_linear_bwd_dw(const float* gzout, int64_t gzout_plane_stride, int32_t ldgzout,
                                const float* zin, int64_t zin_plane_stride, int32_t ldzin, const pcfd_intrans_t* tin_host,
                                float* gw, int32_t ldgw, float* gbias, float* gcvec, int32_t ldgcvec,
                                int32_t cj, int64_t rows, int64_t rows_per_geom, int32_t k, int32_t n,
                                void* workspace, size_t workspace_bytes, void* stream);

/*
 * Segmented max with arg-max: out[s][c] = max_{valid slots j} act(z[s*seg_len + j][c]).
 * Replaces torch.max(dim=1) (models/modules.py:81, :190, :214), PyG global_max_pool (:420) and
 * the max aggregation of PointNetConv (:277-292).  `slots` ([n_seg][seg_len] int32, <0 = empty)
 * masks padded edge slots, NULL = all valid.  Ties -> lowest slot.  Empty segment -> 0, arg -1.
 */
int pcfd_segmax_fwd(const float* z, int32_t ldz, int32_t act, const int32_t* slots,
                    int64_t n_seg, int32_t seg_len, int32_t c,
                    float* out, int32_t ldout, int32_t* arg, void* stream);
/* The same, also writing zsel[s][c] = z[s*seg_len + arg[s][c]][c], the PRE-activation of the selected row (0 for an
 * empty segment): what the sparse backward below needs for act'(.) without a second, dependent gather. */
int pcfd_segmax_fwd_z(const float* z, int32_t ldz, int32_t act, const int32_t* slots,
                      int64_t n_seg, int32_t seg_len, int32_t c,
                      float* out, int32_t ldout, int32_t* arg, float* zsel, int32_t ldzsel, void* stream);
/* gz[s*seg_len + j][c] = (j == arg[s][c]) ? gout[s][c] * act'(z) : 0   (gz fully overwritten) */
int pcfd_segmax_bwd(const float* gout, int32_t ldgout, const int32_t* arg,
                    const float* z, int32_t ldz, int32_t act,
                    int64_t n_seg, int32_t seg_len, int32_t c,
                    float* gz, int32_t ldgz, void* stream);

/*
 * Backward of "last layer of a value-only MLP -> max pool" in sparse form.  Every pooled encoder of the reference
 * ends in  out[s][c] = max_j act(z[s*seg_len + j][c]),  z = T(zin) W^T + b  (PointNetConv max aggregation
 * models/modules.py:286-292, GlobalSetAbstraction :412-423, Branch :184-190, GeometryEncoder :206-214), so the
 * cotangent of z has one non-zero entry per (segment, channel), at row arg[s][c].  Instead of materialising it
 * (pcfd_segmax_bwd) and running the dense dX / dW over all rows, this entry point ACCUMULATES
 *     gw[c][:] += g[s][c] * T(zin)[s*seg_len + arg[s][c]][:],   gbias[c] += g[s][c]        (gw / gbias optional)
 * and OVERWRITES (optional)
 *     gzin[s*seg_len + j][:] = T'(zin) * sum_{c: arg[s][c] == j} g[s][c] * W[c][:]          (zero rows where nothing pooled)
 * with g[s][c] = gout[s][c] * act'(zsel[s][c]) (zsel from pcfd_segmax_fwd_z).  Fixed summation order (deterministic).
 * `tin_host` may carry an activation over all k columns; dropout / escale / partial activation are not supported
 * here (pcfd_pool_layer_bwd_supported returns 0 and the caller takes pcfd_segmax_bwd + the dense layer backward).
 */
int pcfd_pool_layer_bwd_supported(int64_t n_seg, int32_t seg_len, int32_t k, int32_t c, const pcfd_intrans_t* tin_host,
                                  int32_t ldzin);
size_t pcfd_pool_layer_bwd_workspace_bytes(int64_t n_seg, int32_t seg_len, int32_t k, int32_t c);
int pcfd_pool_layer_bwd(const float* gout, int32_t ldgout, const int32_t* arg, const float* zsel, int32_t ldzsel,
                        int32_t act_pool, int64_t n_seg, int32_t seg_len, int32_t c,
                        const float* zin, int32_t ldzin, const pcfd_intrans_t* tin_host, int32_t k,
                        const float* w, int32_t ldw, float* gw, int32_t ldgw, float* gbias,
                        float* gzin, int32_t ldgzin, void* workspace, size_t workspace_bytes, void* stream);
/*
 * Long segments (seg_len > c, global pools over thousands of points): at most c rows of a segment receive gradient,
 * and every layer in front of the pool acts row by row, so their backward may run on those rows alone.  Emits the
 * compacted problem: ids[s][ch] = arg[s][ch] (int64 row inside the segment, 0 for an empty segment) and the cotangent
 * gzc[(s*c + ch)][:] = g[s][ch] * e_ch of compact row (s, ch) (all ldgzc columns written).  A row selected by several
 * channels appears once per channel: the backward is linear in the cotangent, so the contributions add up.
 */
int pcfd_pool_compact(const float* gout, int32_t ldgout, const int32_t* arg, const float* zsel, int32_t ldzsel,
                      int32_t act_pool, int64_t n_seg, int32_t c,
                      int64_t* ids, float* gzc, int32_t ldgzc, void* stream);

/*
 * Farthest point sampling, one CTA per geometry; torch_cluster.fps(pos, batch, ratio) as called at
 * models/modules.py:320 with a deterministic start (first point; upstream default is random).
 * pos [n_geom][n][dims]; m = ceil(ratio*n) samples per geometry; idx_out [n_geom][m] int64 indices
 * into the FLATTENED point array (g*n + i), in selection order.  Squared distances are
 * accumulated left to right in fp32 without FMA contraction; ties -> lowest index (bit-exact
 * against oracle/pyg_restate.py).
 */
int pcfd_fps(const float* pos, int32_t n_geom, int32_t n, int32_t dims, int32_t m,
             int64_t* idx_out, void* stream);
/* Point sets whose coordinates + distances exceed one SM's shared memory (n > ~14k in 3-D) keep their running
 * min-distances in `workspace` (pcfd_fps_workspace_bytes, 0 for the shared-memory / register paths);
 * pcfd_fps == pcfd_fps_ws without a workspace and returns PCFD_ERR_WORKSPACE for such sizes. */
size_t pcfd_fps_workspace_bytes(int32_t n_geom, int32_t n, int32_t dims);
int pcfd_fps_ws(const float* pos, int32_t n_geom, int32_t n, int32_t dims, int32_t m,
                int64_t* idx_out, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Ball query, torch_cluster.radius(x=pos, y=pos[idx], r, batch, batch[idx], K) as called at
 * models/modules.py:321: for each centroid the first k points (ascending index) of the same
 * geometry with squared distance < r*r (strict).  nbr [n_geom*m][k] int32 flattened point indices,
 * -1 padded; count [n_geom*m].
 */
int pcfd_ball_query(const float* pos, const int64_t* centroid_idx, int32_t n_geom, int32_t n, int32_t dims,
                    int32_t m, float r, int32_t k, int32_t* nbr, int32_t* count, void* stream);

/*
 * Edge slots of one SetAbstraction layer with PyG PointNetConv's self-loop rule on the flattened
 * bipartite graph (models/modules.py:322-323): drop neighbours whose flattened point index equals
 * the flattened centroid index, then append source point i -> centroid i.  slots [m_total][k+1].
 */
int pcfd_sa_edges(const int32_t* nbr, int64_t m_total, int32_t k, int64_t n_points_total,
                  int32_t* slots, void* stream);
/*
 * Set-abstraction geometry of a batch from a per-geometry cache (the dataset is sampled once,
 * dataset/foam_dataset.py:159-161, so FPS centroids and ball-query neighbourhoods of a geometry never change).
 * idx_local [n_geom][m] (int64) and nbr_local [n_geom][m][k] (int32, -1 = empty) hold indices local to their geometry
 * (0 .. n-1) for the geometries of THIS batch, in batch order; the call rebases them to the flattened numbering of the
 * batch and builds the edge slots with the rule of pcfd_sa_edges: idx [n_geom*m], slots [n_geom*m][k+1].
 */
int pcfd_sa_cached_geometry(const int64_t* idx_local, const int32_t* nbr_local, int32_t n_geom, int32_t m, int32_t k,
                            int32_t n, int64_t* idx, int32_t* slots, void* stream);
/*
 * Edge features (PointConvNext.message, models/modules.py:286-292):
 *   ein[(i,s)] = [ x[j][0:f_in], pos[j] - pos[centroid_i] / r ],  j = slots[i][s]; zeros if empty.
 * ein rows are written 16 bytes at a time: ein 16-byte aligned, ldein a multiple of 4 (>= f_in + dims); the pad columns
 * of a row are set to zero.
 */
int pcfd_sa_gather(const float* x, int32_t ldx, int32_t f_in, const float* pos, int32_t dims,
                   const int64_t* centroid_idx, const int32_t* slots, int64_t m_total, int32_t kp, float r,
                   float* ein, int32_t ldein, void* stream);
/* gx[j][0:f_in] += gein[(i,s)][0:f_in] for every valid slot (atomic adds; gx is accumulated). */
int pcfd_sa_scatter_bwd(const float* gein, int32_t ldgein, const int32_t* slots, int64_t m_total, int32_t kp,
                        int32_t f_in, float* gx, int32_t ldgx, void* stream);

/*
 * FoamData indexing folded into one gather (dataset/foam_data.py:36-61: label -> column slice,
 * sub-domain -> torch.gather):  out[g*out_rows_per_geom + out_row_offset + i][out_col_offset + c]
 *   = data[g][row_ids[g][i]][cols[c]].   row_ids NULL -> rows first_row .. first_row+n_sel-1.
 * cols_host: up to 32 column indices, read on the host at call time; NULL = columns 0..n_cols-1
 * (any width: a strided 2-D copy).
 */
int pcfd_gather_cols(const float* data, int32_t n_geom, int64_t n_rows, int32_t f,
                     const int64_t* row_ids, int64_t first_row, int64_t n_sel,
                     const int32_t* cols_host, int32_t n_cols,
                     float* out, int32_t ldout, int64_t out_rows_per_geom, int64_t out_row_offset,
                     int32_t out_col_offset, void* stream);

/* Seeds the input jet of the per-point path: plane 0 = coordinates, plane k = e_k, others 0.
 * (enable_internal_autograd, models/model_base.py:56-66: the points become the autograd leaf.) */
int pcfd_seed_jet(const float* data, int32_t n_geom, int64_t n_rows, int32_t f,
                  const int64_t* row_ids, int64_t n_sel, const int32_t* coord_cols_host, int32_t dims,
                  int32_t cj, float* zout, int64_t plane_stride, int32_t ldz, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused residual assembly + losses + their gradient (models/losses.py:149-319, :10-20, :55-56;
 * models/model_base.py:191-212).
 * ------------------------------------------------------------------------------------------ */
enum { PCFD_LOSS_MANUFACTURED = 0, PCFD_LOSS_FIXED = 1, PCFD_LOSS_VARIABLE = 2 };
enum { PCFD_LAP_REFERENCE = 0, PCFD_LAP_TRUE = 1 };
#define PCFD_LOSS_OUT_FLOATS 48
/* layout of `out`: [0..15] unscaled loss terms in the reference's order
 * (continuity, momentum xD, boundary U xD, boundary p, [obs U xD, obs p]); [16..31] the same
 * times the loss weights; [32] their sum; [33..35] MAE of U per component, [36] MAE of p
 * (calculate_errors, models/model_base.py:168-180); [37] number of loss terms. */

typedef struct pcfd_residual_params {
  int32_t dims, loss_kind, lap_mode, enable_data_loss;
  float nu, d, f;                 /* scalar Darcy / Forchheimer coefficients (manufactured, fixed) */
  float c_std[3], u_std[3], u_mean[3], p_std, p_mean;
  float d_min[3], d_range[3], f_min[3], f_range[3];
  int32_t col_u[3], col_p, col_zone, col_d[3], col_f[3]; /* column indices into data */
  float weights[16];              /* FixedLossScaler weights (all 1 if the model has no scaler) */
} pcfd_residual_params_t;

size_t pcfd_residual_workspace_bytes(int32_t n_geom, int64_t ni, int64_t nb, int64_t no);
/*
 * y_int: jets of the model output at the internal points [cj][n_geom*ni][ldy] (cj = 1+D for
 * PCFD_LAP_REFERENCE, 1+2D for PCFD_LAP_TRUE); y_bnd: values at the boundary points
 * [n_geom*nb][ldy].  Writes gy_int / gy_bnd (same shapes, overwritten) = d(sum of scaled
 * losses)/dy, and `out` (PCFD_LOSS_OUT_FLOATS floats).  Row ids index data[g] (targets) and the
 * prediction order (internal rows first, then boundary rows), as in the reference.
 */
int pcfd_residual_loss(const float* data, int32_t n_geom, int64_t n_rows, int32_t f,
                       const int64_t* internal_ids, int64_t ni, const int64_t* boundary_ids, int64_t nb,
                       const int64_t* obs_ids, int64_t no,
                       const float* y_int, int64_t y_plane_stride, const float* y_bnd, int32_t ldy,
                       const pcfd_residual_params_t* prm_host,
                       float* gy_int, float* gy_bnd, float* out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* pcfd_residual_loss with DEVICE-resident loss weights (`weights_dev`, one float per loss term; NULL = prm->weights):
 * the weights of an adaptive scaler change every step without re-recording a captured graph.
 * `visc_extra` (optional, [n_geom*ni][D]) is added to the Laplacian row sums of the momentum residual and `gvisc`
 * (optional, same shape) receives d loss / d visc_extra: the cross-point terms of vanilla PIPN (coupling.py). */
int pcfd_residual_loss_w(const float* data, int32_t n_geom, int64_t n_rows, int32_t f,
                         const int64_t* internal_ids, int64_t ni, const int64_t* boundary_ids, int64_t nb,
                         const int64_t* obs_ids, int64_t no,
                         const float* y_int, int64_t y_plane_stride, const float* y_bnd, int32_t ldy,
                         const pcfd_residual_params_t* prm_host, const float* weights_dev,
                         const float* visc_extra, float* gvisc,
                         float* gy_int, float* gy_bnd, float* out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* The residual stage of a training step in ONE kernel launch (north-star (c): continuity, momentum with the
 * Darcy-Forchheimer terms, boundary and observation MSE, weighting, reduction and d loss / d jet): same arguments and
 * results as pcfd_residual_loss_w.  Blocks take the internal, boundary and observation points by role; the block that
 * finishes last reduces the per-block partial sums in a fixed order.  The value plane of gy is accumulated with
 * atomics, so it is zeroed first by one cudaMemsetAsync on `stream` (one memset node when gy_int directly follows
 * gy_bnd in memory).  `ticket`: a device int32 that is zero before the first call and is left at zero by every call;
 * the caller owns it and must not share it between streams.  Replaces the same reference code as pcfd_residual_loss
 * (models/losses.py:149-319, models/model_base.py:191-218). */
int pcfd_residual_step(const float* data, int32_t n_geom, int64_t n_rows, int32_t f,
                       const int64_t* internal_ids, int64_t ni, const int64_t* boundary_ids, int64_t nb,
                       const int64_t* obs_ids, int64_t no,
                       const float* y_int, int64_t y_plane_stride, const float* y_bnd, int32_t ldy,
                       const pcfd_residual_params_t* prm_host, const float* weights_dev,
                       const float* visc_extra, float* gvisc,
                       float* gy_int, float* gy_bnd, float* out, int32_t* ticket,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Residual fields at inference, predict_step with verbose_predict (models/model_base.py:233-252):
 * fields [n_geom*ni][D+1] = cat([momentum residual (D), divergence]) at the internal points. */
int pcfd_residual_fields(const float* data, int32_t n_geom, int64_t n_rows, int32_t f,
                         const int64_t* internal_ids, int64_t ni,
                         const float* y_int, int64_t y_plane_stride, int32_t ldy,
                         const pcfd_residual_params_t* prm_host, float* fields, void* stream);

/*
 * The loss modules on EXPLICIT tensors, for callers that hold the derivative tensors themselves (the reference's
 * predict_step, models/model_base.py:241-246, and evaluation code): ContinuityLoss[Standardized].func
 * (models/losses.py:154-156, 177-182) -> div [rows]; MomentumLoss{Manufactured,Fixed,Variable}.func (:209-217,
 * :256-266, :301-311) -> momentum [rows][D].  Dense row-major inputs: u [rows][D], jac / lap [rows][D][D]
 * (jac[i][j] = dU_i/dx_j, lap[i][j] = d2U_i/dx_j2), p_grad [rows][D], zone [rows] (cellToRegion), and per kind:
 * VARIABLE dcoef / fcoef [rows][D] (normalised d, f), MANUFACTURED fcoef [rows][D] (the forcing term, may be NULL).
 * Either output may be NULL.  Scalers / constants come from prm_host (column indices and weights are ignored).
 */
int pcfd_residual_eval(const float* u, const float* jac, const float* lap, const float* p_grad, const float* zone,
                       const float* dcoef, const float* fcoef, int64_t rows, const pcfd_residual_params_t* prm_host,
                       float* momentum, float* div, void* stream);
/* out[c] = mean_r x[r][c]^2: vector_loss(res, 0, mse_loss) / mse_loss(res, 0) of the modules' forward (models/losses.py:10-20). */
int pcfd_mean_squares(const float* x, int64_t rows, int32_t cols, float* out, void* stream);

/* RelobraloScaler.forward (models/losses.py:93-124) on the device: `losses` = the n unscaled loss terms of this
 * step (out[0..n) of a residual pass), the three buffers are the module's registered buffers, `step` a device
 * counter (the reference's global_step, incremented here), `batch_size` what the reference reads from
 * trainer.train_dataloader.batch_size.  Writes the n weights the residual pass applies (all 1 at step 0).
 * rho ~ Bernoulli(beta) is drawn from a hash of (seed, step). */
int pcfd_relobralo_update(const float* losses, int32_t n, float* init_losses, float* prev_losses, float* lambda_ema,
                          int64_t* step, int32_t batch_size, float alpha, float beta, float tau, float eps,
                          uint64_t seed, float* weights_out, void* stream);

/* Data-parallel optimizer tail in one kernel over NVLink / NVSwitch peer memory (csrc/dp_adam.cu): gradient
 * reduce-scatter + Adam on this rank's slice + all-gather of the new parameters, replacing ncclAllReduce(flat gradient)
 * followed by pcfd_adam_step (reference: Lightning DDP + torch.optim.Adam, common/training.py:66-71,83).
 * `peers`: for every rank r < world the address, IN THIS PROCESS, of rank r's flat gradient, flat parameters and flag
 * array (symmetric memory: same sizes on every rank, peer-mapped; buffers padded to a multiple of 4 floats and 16-byte
 * aligned), plus the multicast addresses of the gradient and parameter buffers (both NULL: peer loads / stores instead
 * of multimem.ld_reduce / multimem.st).  The flag array holds pcfd_dp_flags_len() uint32, zero before the first call.
 * `epoch`: 4 int32 of local device memory, zero before the first call ([2] != 0 afterwards = a peer did not arrive
 * within ~2 s).  Every rank must call this once per step with the same n and hyper-parameters; exp_avg / exp_avg_sq
 * are full-size local buffers of which only this rank's slice is used.  grad_scale = 1 / (world * micro-batches). */
typedef struct {
  void* grad[16];
  void* param[16];
  void* flags[16];
  void* grad_mc;
  void* param_mc;
  int32_t rank, world;
} pcfd_dp_peers_t;
int32_t pcfd_dp_flags_len(void);
int pcfd_dp_adam_step(const pcfd_dp_peers_t* peers, float* exp_avg, float* exp_avg_sq, int64_t* step, const float* lr,
                      float beta1, float beta2, float eps, float grad_scale, int64_t n, int32_t* epoch, void* stream);

/* torch.optim.Adam (no weight decay, no amsgrad; models/pipn/pipn_foam.py:102-105 configure_optimizers) on flat
 * fp32 buffers: step += 1; m, v, param updated in one pass.  `step` (int64) and `lr` (float) are device scalars so
 * that the launch is graph-capturable; `grad_scale` multiplies the gradient first (1/world after an all-reduce). */
int pcfd_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t* step,
                   const float* lr, float beta1, float beta2, float eps, float grad_scale, int64_t n, void* stream);

/* out[i] = 0 for i < n (graph-capturable memset of gradient buffers) */
int pcfd_zero(float* p, int64_t n, void* stream);
/* *seed_dev = mix(*seed_dev) : advances the dropout seed once per step without a host round trip */
int pcfd_advance_seed(uint64_t* seed_dev, void* stream);

/*
 * Batch ingestion on the device.  `data` is the dataset tensor [n_geom][n_points][f] with the
 * internal points of every geometry in rows [0, n_internal) and its boundary points behind them (the row order
 * FoamDataset.load_case produces, dataset/foam_dataset.py:424-433).
 *
 * pcfd_sdf_feature: FoamDataset.add_sdf (dataset/foam_dataset.py:360-380).  data[g][p][sdf_col] = distance of point p
 * to the nearest boundary point of geometry g (scipy cdist + min), divided by the largest such distance of the
 * geometry, times (0.5 - data[g][p][region_col]) * 2 for the internal points (region_col < 0: no sign).  Coordinates
 * are columns pos_col .. pos_col+dims-1, multiplied by coord_scale[dims] first (device pointer or NULL: the `range` /
 * `std` of the coordinate scaler, whose inverse_transform the reference applies before measuring distances).
 * scratch: pcfd_sdf_scratch_bytes(n_geom, n_points) bytes.
 */
int pcfd_sdf_feature(float* data, int32_t n_geom, int32_t n_points, int32_t f, int32_t n_internal, int32_t pos_col,
                     int32_t dims, int32_t region_col, int32_t sdf_col, const float* coord_scale, float* scratch,
                     void* stream);
size_t pcfd_sdf_scratch_bytes(int32_t n_geom, int32_t n_points);
/* FoamDataset.add_boundary_id (dataset/foam_dataset.py:382-395): columns col0 .. col0+n_classes-1 become zero for the
 * internal rows and the one-hot class of the boundary rows; boundary_class [n_geom][n_points - n_internal] holds the
 * position of each boundary row's patch name in the sorted list of patch names (what OneHotEncoder assigns). */
int pcfd_boundary_one_hot(float* data, int32_t n_geom, int32_t n_points, int32_t f, int32_t n_internal,
                          const int32_t* boundary_class, int32_t n_classes, int32_t col0, void* stream);
/* collate_fn (dataset/foam_dataset.py:83-90) for a dataset resident in HBM: dst[i] = src[ids[i]] for n_ids blocks of
 * block_bytes (a multiple of 4; 16-byte copies when everything is 16-byte aligned).  Used for the data tensor and for
 * every sub-domain's row ids. */
int pcfd_gather_blocks(const void* src, int64_t block_bytes, const int64_t* ids, int64_t n_ids, void* dst, void* stream);
/* The same for up to PCFD_GATHER_MAX_TENSORS tensors of one dataset in ONE launch (the data tensor and the row ids of
 * every sub-domain): dst[t][i] = src[t][ids[i]].  The three arrays are host arrays read at call time; ids is a device
 * array; a block size of 0 (an empty sub-domain) is skipped. */
#define PCFD_GATHER_MAX_TENSORS 16
int pcfd_gather_blocks_multi(const void* const* src_host, void* const* dst_host, const int64_t* block_bytes_host,
                             int32_t n_tensors, const int64_t* ids, int64_t n_ids, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCFD_H_ */
