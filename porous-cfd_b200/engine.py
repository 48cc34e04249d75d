"""Host-side executor of the fused physics-informed training step.

It replaces the body of the reference's `PorousPinnBase.training_step`
(models/model_base.py:182-218: forward, 1 + D + D*D + 1 reverse sweeps, residual losses, double
backward) by ONE forward-mode jet pass through the shared per-point MLP stack, one fused residual
kernel and one reverse pass, all of it C-ABI calls into libpcfd_sm100.so on the current CUDA
stream (graph-capturable).  Python only sequences the calls and owns the buffers.

Structure of a step:
  encode      per-geometry constants: pooled global feature (MLP+max or set-abstraction stack),
              branch embedding; folded into a per-geometry vector `cvec` added by the first layer
              that consumes the concat [per-point features, global feature]
  point chain jet pass on the internal points (cj = 1+D or 1+2D channels) and value pass on the
              boundary points (cj = 1) through the same weights
  residual    continuity / momentum / boundary / observation losses and d loss / d jet
  backward    reverse chain (dX, dW), reverse encode
"""
from __future__ import annotations

import math
import os
import weakref
from dataclasses import dataclass, field
from typing import Optional

import torch
from torch import Tensor, nn

from . import _lib, ops
from .ops import Jet


@dataclass
class ChainLayer:
    weight: nn.Parameter
    bias: Optional[nn.Parameter]
    col_lo: int
    k: int
    n: int
    act: Optional[str] = None        # transform applied to this layer's INPUT while it is loaded
    act_cols: int = 0
    drop_p: float = 0.0
    escale: bool = False
    cvec_key: Optional[str] = None   # per-geometry constant (bias folded into it) added to channel 0


def mlp_chain(linears, act: str, last_activation: bool, dropout=None, first_act: Optional[str] = None,
              first_drop: float = 0.0):
    """Chain of a Linear->act(->Dropout) stack.  Activation/dropout of layer i are applied by
    layer i+1 on load; returns (layers, pending) where pending = (act, drop_p) still to be applied
    to the output of the last layer by whoever consumes it."""
    layers = []
    pend_act, pend_drop = first_act, first_drop
    n_lin = len(linears)
    for i, lin in enumerate(linears):
        layers.append(ChainLayer(lin.weight, lin.bias, 0, lin.in_features, lin.out_features, act=pend_act,
                                 drop_p=pend_drop))
        pend_act = act if (i < n_lin - 1 or last_activation) else None
        pend_drop = float(dropout[i]) if dropout is not None else 0.0
    return layers, (pend_act, pend_drop)


class StepContext:
    """Buffers shared by the calls of one executor: gradient views, scratch workspaces, dropout seed, streams."""

    def __init__(self, device):
        self.device = device
        # fork independent work to side streams (off: one stream, for per-kernel timing / ordered ncu launch lists)
        self.overlap = os.environ.get('PCFD_NO_OVERLAP', '0') != '1'
        # One scratch workspace PER STREAM (split partials of the dW kernels, residual partial sums): calls on one stream
        # are ordered, so a stream's workspace can never be in use by two kernels at once, whatever runs beside it.
        self.workspaces: dict = {}
        self.retired: list = []
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=device)
        self.training = False
        self.grads: dict[int, Tensor] = {}
        self.salt = 0
        self.side_stream = None
        self.streams: dict = {}
        self.keep: list = []          # tensors produced on one stream and read on another: alive until the next step begins

    def stream(self, name: str) -> torch.cuda.Stream:
        st = self.streams.get(name)
        if st is None:
            st = self.streams[name] = torch.cuda.Stream(device=self.device)
        return st

    def need_workspace(self, nbytes: int) -> Tensor:
        """The current stream's workspace, at least `nbytes` large."""
        key = torch.cuda.current_stream().cuda_stream
        ws = self.workspaces.get(key)
        if ws is None or ws.numel() < nbytes:
            # a captured step graph has the address of the workspace it was recorded with baked in: superseded buffers
            # stay allocated (they are a few MB) so that replaying an older graph never writes into freed memory
            if ws is not None:
                self.retired.append(ws)
            ws = self.workspaces[key] = torch.empty(max(int(nbytes * 1.25) + 256, 1 << 20), dtype=torch.uint8, device=self.device)
        return ws

    @property
    def workspace(self) -> Tensor:
        """The current stream's workspace as sized by the last need_workspace() on this stream."""
        return self.need_workspace(0)

    def grad(self, p: Optional[nn.Parameter]) -> Optional[Tensor]:
        return None if p is None else self.grads[id(p)]


def _tin(ctx: StepContext, layer: ChainLayer, escale: Optional[Tensor], salt: int):
    drop = layer.drop_p if ctx.training else 0.0
    if layer.act is None and drop == 0.0 and not layer.escale:
        return None
    return ops.make_intrans(layer.act, layer.act_cols, escale if layer.escale else None, drop, ctx.seed_dev, salt)


def chain_forward(ctx: StepContext, layers, z0: Jet, rows_per_geom: int, escale: Optional[Tensor] = None,
                  cvecs: Optional[dict] = None, salt_base: int = 0):
    zs = [z0]
    for i, L in enumerate(layers):
        tin = _tin(ctx, L, escale, salt_base + i)
        cvec = cvecs[L.cvec_key] if L.cvec_key is not None else None
        bias = None if L.cvec_key is not None else L.bias
        zs.append(ops.jet_linear_fwd(zs[-1], tin, L.weight, L.col_lo, L.k, bias, cvec, rows_per_geom, L.n))
    return zs


def chain_backward(ctx: StepContext, layers, zs, gz: Jet, rows_per_geom: int, escale: Optional[Tensor] = None,
                   gescale: Optional[Tensor] = None, gcvecs: Optional[dict] = None, need_input_grad: bool = False,
                   salt_base: int = 0, side: Optional[torch.cuda.Stream] = None) -> Optional[Jet]:
    """Reverse pass of a chain.  With `side`, the weight-gradient work of every layer (dW kernel + its small,
    latency-bound finish kernel) is issued on that stream while the main stream goes on with dX: the two big
    kernels still share the SMs one after the other, but the finish kernels no longer sit on the critical path.
    The caller joins the streams; every stream has its own workspace (StepContext.need_workspace)."""
    main = torch.cuda.current_stream()
    for i in range(len(layers) - 1, -1, -1):
        L = layers[i]
        zin = zs[i]
        tin = _tin(ctx, L, escale, salt_base + i)
        nbytes = ops.dw_workspace_bytes(zin.cj, zin.rows, rows_per_geom, L.k, L.n)
        gbias = ctx.grad(L.bias) if L.cvec_key is None else None
        gcvec = gcvecs[L.cvec_key] if L.cvec_key is not None else None
        if side is not None:
            side.wait_stream(main)
            gz.t.record_stream(side)
            ctx.keep += [gz, zin]
            with torch.cuda.stream(side):
                ops.jet_linear_bwd_dw(gz, zin, tin, ctx.grad(L.weight), L.col_lo, gbias, gcvec, rows_per_geom, L.k, L.n,
                                      ctx.need_workspace(nbytes))
        else:
            ops.jet_linear_bwd_dw(gz, zin, tin, ctx.grad(L.weight), L.col_lo, gbias, gcvec, rows_per_geom, L.k, L.n,
                                  ctx.need_workspace(nbytes))
        if i > 0 or need_input_grad:
            gz = ops.jet_linear_bwd_dx(gz, L.weight, L.col_lo, zin, tin, gescale if L.escale else None, rows_per_geom,
                                       L.k, L.n)
    return gz if need_input_grad else None


def pool_backward(ctx: StepContext, layers, zs, gout: Tensor, ldgout: int, arg: Tensor, zsel: Tensor, act_pool, n_seg: int,
                  seg_len: int,
                  rows_per_geom: int = 0, need_input_grad: bool = False,
                  side: Optional[torch.cuda.Stream] = None) -> Optional[Jet]:
    """Reverse pass of `chain(layers) -> max pool over segments of seg_len rows`.  The cotangent of the last layer's
    output has one non-zero entry per (segment, channel), so instead of materialising it (segmax_bwd) and running the
    dense dX / dW over every row:
      * short / medium segments (set-abstraction neighbourhoods, global set abstraction): sparse last-layer backward
        (ops.pool_layer_bwd), then the ordinary dense reverse pass of the layers in front of it;
      * long segments with fewer channels than rows (PI-GANO geometry / branch encoders): only the <= C selected rows of a
        segment carry gradient and every layer acts row by row, so the whole reverse pass runs on the compacted rows;
      * anything else (dropout / branch scaling on the last layer): the dense form.
    Measured on B200 (scripts/bench_pool.py): the sparse kernels are CUDA-core kernels with ~10 instructions per gathered
    16-byte chunk, the dense form runs on the tensor cores; with K + 1 = 17 slots per neighbourhood (config 2) the dense
    form is as fast (174 vs 183 us for 272 000 edge rows 64 -> 128), from ~48 slots on (windbreaks: 65) and for thin inputs
    (k <= 16, manufactured set abstraction) the sparse form wins (207 vs 290 us; 34 vs 226 us), so that is the switch.
    PCFD_POOL_SPARSE=0 forces the dense form, =2 the sparse form wherever it is supported (tests compare them)."""
    L = layers[-1]
    zlast, zin = zs[-1], zs[-2]
    c, k = L.n, L.k
    tin = _tin(ctx, L, None, 0)
    plain = L.cvec_key is None and not L.escale and L.col_lo == 0 and (not ctx.training or L.drop_p == 0.0)
    mode = os.environ.get('PCFD_POOL_SPARSE', '1')
    worth = mode == '2' or k <= 16 or (seg_len >= 48 and n_seg >= 512)
    if mode != '0' and worth and plain and ops.pool_layer_bwd_supported(n_seg, seg_len, k, c, tin, zin.ld):
        need_gzin = len(layers) > 1 or need_input_grad
        ws_bytes = ops.pool_layer_bwd_workspace_bytes(n_seg, seg_len, k, c)
        main = torch.cuda.current_stream()
        if side is not None:
            side.wait_stream(main)
            gout.record_stream(side)
            ctx.keep.append(gout)
            with torch.cuda.stream(side):
                ops.pool_layer_bwd(gout, ldgout, arg, zsel, act_pool, n_seg, seg_len, c, zin, tin, k, L.weight,
                                   ctx.grad(L.weight), ctx.grad(L.bias), False, ctx.need_workspace(ws_bytes))
        else:
            ops.pool_layer_bwd(gout, ldgout, arg, zsel, act_pool, n_seg, seg_len, c, zin, tin, k, L.weight,
                               ctx.grad(L.weight), ctx.grad(L.bias), False, ctx.need_workspace(ws_bytes))
        if not need_gzin:
            return None
        gzin = ops.pool_layer_bwd(gout, ldgout, arg, zsel, act_pool, n_seg, seg_len, c, zin, tin, k, L.weight,
                                  None, None, True, None)
        if len(layers) == 1:
            return gzin
        return chain_backward(ctx, layers[:-1], zs[:-1], gzin, rows_per_geom, need_input_grad=need_input_grad, side=side)
    no_drop = all((not ctx.training or l.drop_p == 0.0) and not l.escale and l.cvec_key is None for l in layers)
    if mode != '0' and seg_len > c and not need_input_grad and no_drop:
        ids, gzc = ops.pool_compact(gout, ldgout, arg, zsel, act_pool, n_seg, c)
        zs_c = []
        for zj in zs[:-1]:
            zc = Jet.empty(1, n_seg * c, zj.width, zj.t.device)
            ops.gather_cols(zj.t, n_seg, seg_len, zj.ld, ids, 0, c, list(range(zj.width)), zc.t, zc.ld, c)
            zs_c.append(zc)
        return chain_backward(ctx, layers, zs_c, gzc, 0, need_input_grad=False, side=side)
    gz = ops.segmax_bwd(gout, ldgout, arg, zlast.t[0], act_pool, n_seg, seg_len, c)
    return chain_backward(ctx, layers, zs, Jet(gz, c), rows_per_geom, need_input_grad=need_input_grad, side=side)


# ------------------------------------------------------------------------------------------------
# set-abstraction stack (models/modules.py:94-98, 295-325, 403-423, 483-527)
# ------------------------------------------------------------------------------------------------

@dataclass
class SALevel:
    ratio: float
    radius: float
    layers: list
    act: str
    max_neighbors: int


@dataclass
class SAStack:
    levels: list
    global_layers: Optional[list]
    act: str
    out_features: int = 0


def sa_geometry_levels(stack: SAStack, pos0: Tensor) -> list:
    """FPS centroids, ball-query neighbours and centroid positions of every level: per level {'idx' (B, m) int64 and
    'nbr' (B*m, K) int32, both flattened over the batch, 'newpos' (B, m, D), 'n' points per geometry at this level}."""
    b, n, d = pos0.shape
    pos, out = pos0, []
    for lvl in stack.levels:
        idx = ops.fps(pos, lvl.ratio)
        m = idx.shape[1]
        nbr, _ = ops.ball_query(pos, idx, lvl.radius, lvl.max_neighbors)
        newpos = torch.empty((b, m, d), dtype=torch.float32, device=pos.device)
        ops.gather_cols(pos, 1, b * n, d, idx, 0, b * m, list(range(d)), newpos, d, b * m)
        out.append({'idx': idx, 'nbr': nbr, 'newpos': newpos, 'n': n})
        pos, n = newpos, m
    return out


def sa_geometry(stack: SAStack, pos0: Tensor) -> list:
    """The part of the set-abstraction stack that depends on the point positions only (no weights, no features):
    per level the FPS centroids, the radius neighbourhoods as edge slots, and the centroid positions.  A training loop
    that knows its next batch can run this for batch t+1 beside the step of batch t (PinnExecutor.geometry)."""
    b = pos0.shape[0]
    geo = []
    for lv in sa_geometry_levels(stack, pos0):
        slots = ops.sa_edges(lv['nbr'], b * lv['n'])
        geo.append({'idx': lv['idx'], 'slots': slots, 'newpos': lv['newpos']})
    return geo


def sa_forward(ctx: StepContext, stack: SAStack, x0: Tensor, ldx0: int, f0: int, pos0: Tensor, geo: Optional[list] = None):
    """x0 [B*n0, ldx0] features, pos0 (B, n0, D) -> (g [B, ld] pooled feature, saved state).  `geo`: the output of
    sa_geometry(stack, pos0) when it was computed ahead of the step."""
    b, n, d = pos0.shape
    x, ldx, f, pos = x0, ldx0, f0, pos0
    saved = {'levels': [], 'b': b, 'd': d}
    if geo is None:
        geo = sa_geometry(stack, pos0)
    for lvl, gl in zip(stack.levels, geo):
        idx, slots, newpos = gl['idx'], gl['slots'], gl['newpos']
        m = idx.shape[1]
        ein = ops.sa_gather(x, ldx, f, pos, idx, slots, lvl.radius)
        zs = chain_forward(ctx, lvl.layers, Jet(ein, f + d), 0)
        c = lvl.layers[-1].n
        out, arg, zsel = ops.segmax_fwd_z(zs[-1].t[0], lvl.act, slots, b * m, slots.shape[1], c)
        saved['levels'].append({'slots': slots, 'zs': zs, 'arg': arg, 'zsel': zsel, 'm': m, 'n': n, 'f_in': f, 'ldx': ldx, 'c': c})
        x, ldx, f, pos, n = out, out.stride(0), c, newpos, m
    if stack.global_layers is None:
        raise NotImplementedError('a set-abstraction stack without a final GlobalSetAbstraction is not used by any '
                                  'in-scope model')
    gin = torch.empty((1, b * n, ops.round4(f + d)), dtype=torch.float32, device=pos.device)
    ops.gather_cols(x, 1, b * n, ldx, None, 0, b * n, list(range(f)), gin, gin.stride(1), b * n, 0, 0)
    ops.gather_cols(pos, 1, b * n, d, None, 0, b * n, list(range(d)), gin, gin.stride(1), b * n, 0, f)
    zs = chain_forward(ctx, stack.global_layers, Jet(gin, f + d), n)
    e = stack.global_layers[-1].n
    g, arg, zsel = ops.segmax_fwd_z(zs[-1].t[0], stack.act, None, b, n, e)
    saved.update({'g_zs': zs, 'g_arg': arg, 'g_zsel': zsel, 'g_n': n, 'g_f': f, 'e': e})
    return g, saved


def sa_backward(ctx: StepContext, stack: SAStack, saved: dict, gg: Tensor, ldgg: int, side=None) -> None:
    b, n, e = saved['b'], saved['g_n'], saved['e']
    zs = saved['g_zs']
    need = len(stack.levels) > 0
    gin = pool_backward(ctx, stack.global_layers, zs, gg, ldgg, saved['g_arg'], saved['g_zsel'], stack.act, b, n, rows_per_geom=n,
                        need_input_grad=need, side=side)
    if not need:
        return
    gx, ldgx = gin.t[0], gin.ld
    for li in range(len(stack.levels) - 1, -1, -1):
        lvl, sv = stack.levels[li], saved['levels'][li]
        m_total = sv['slots'].shape[0]
        zs = sv['zs']
        gein = pool_backward(ctx, lvl.layers, zs, gx, ldgx, sv['arg'], sv['zsel'], lvl.act, m_total, sv['slots'].shape[1],
                             need_input_grad=(li > 0), side=side)
        if li > 0:
            gprev = torch.empty((b * sv['n'], sv['ldx']), dtype=torch.float32, device=gx.device)
            ops.zero_(gprev)
            ops.sa_scatter_bwd(gein.t[0], gein.ld, sv['slots'], sv['f_in'], gprev, sv['ldx'])
            gx, ldgx = gprev, sv['ldx']
            if 'debug' in saved:
                saved['debug'][li] = (gprev, gein)


# ------------------------------------------------------------------------------------------------
# the step executor
# ------------------------------------------------------------------------------------------------

@dataclass
class StepResult:
    out: Tensor                # PCFD_LOSS_OUT_FLOATS device floats (see include/pcfd.h)
    n_terms: int
    y_int: Jet = None
    y_bnd: Jet = None

    @property
    def loss(self) -> Tensor:
        return self.out[32]

    @property
    def losses(self) -> Tensor:
        return self.out[16:16 + self.n_terms]

    @property
    def unscaled(self) -> Tensor:
        return self.out[0:self.n_terms]


class GraphedStep:
    """One captured CUDA graph of `PinnExecutor.step` for a fixed input signature.  Inputs are copied into static
    device buffers (device-to-device, a few microseconds), the ~130 kernels of the step replay as one launch, and the
    results live in static buffers that stay valid until the next replay.

    With `pipeline=True` (models with a set-abstraction encoder) the graph holds a second, independent branch: the
    geometry (FPS, ball query, edge slots) of the NEXT batch, which depends on positions only.  FPS is a sequential
    chain on one CTA per geometry at the very head of the step; computed one step ahead it runs beside the previous
    step's GEMMs instead.  Every replay still does one full step worth of work (geometry of one batch + the rest of
    another).  `run(..., next_data, next_domain)` names the batch the following call will bring; a call whose batch was
    not announced computes its geometry in line first."""

    def __init__(self, ex: 'PinnExecutor', data: Tensor, labels: dict, domain: dict, laplacian: str, pipeline: bool = False,
                 geo: Optional[list] = None):
        self.ex = weakref.proxy(ex)       # the executor owns this object (executor._graphs): no cycle
        self.data = torch.empty_like(data)
        self.domain = {k: torch.empty_like(v) for k, v in domain.items()}
        self.labels = labels
        self.cached = geo is not None          # geometry handed in with every batch (DeviceFoamDataset geometry cache)
        self.pipeline = pipeline and ex.uses_geometry() and not self.cached
        self.load(data, domain)
        self.geo = None
        self.expected = None          # (data, version, boundary ids, version) of the batch whose geometry the static buffers hold
        if self.cached:
            self.geo = [{k: v.clone() for k, v in lv.items()} for lv in geo]      # static buffers the graph reads
        if self.pipeline:
            # the only input of the geometry branch: the next batch's sampled positions, gathered into a static buffer
            self.pos_next = ex.geometry_positions(self.data, labels, self.domain)
            self.geo = ex.geometry(self.data, labels, self.domain)          # static buffers, filled for this first batch
            self.geo_side = torch.cuda.Stream(device=data.device)
        torch.cuda.current_stream().synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # Two precautions around the capture.  (1) Python's cyclic collector must not run inside it: a dead model of an
        # earlier run (model <-> executor <-> GraphedStep is a reference cycle) owns a CUDAGraph and its memory pool, and
        # collecting it DURING this capture frees device memory inside the capture -> cudaErrorStreamCaptureInvalidated
        # (this torch no longer collects in torch.cuda.graph.__enter__; seen in bench.py after its configs block).  So:
        # collect now, keep the collector off until the capture has ended.  (2) thread_local error mode: only this thread's
        # calls are checked against the capture (a DataLoader's pin-memory thread may allocate meanwhile).
        import gc
        gc.collect()
        gc_was_enabled = gc.isenabled()
        gc.disable()
        try:
            with torch.cuda.graph(self.graph, capture_error_mode='thread_local'):
                if self.pipeline:
                    main = torch.cuda.current_stream()
                    self.geo_side.wait_stream(main)
                    with torch.cuda.stream(self.geo_side):
                        nxt = ex.geometry(None, labels, None, pos=self.pos_next)
                    self.result = ex.step(self.data, labels, self.domain, laplacian, geo=self.geo)
                    main.wait_stream(self.geo_side)
                    for cur, new in zip(self.geo, nxt):                          # after the step's last use of the old ones
                        for k in cur:
                            cur[k].copy_(new[k])
                else:
                    self.result = ex.step(self.data, labels, self.domain, laplacian, geo=self.geo)

        finally:
            if gc_was_enabled:
                gc.enable()

    def _announced(self, data: Tensor, domain: dict) -> bool:
        """Are these the tensors the previous call announced, unmodified since?  The announced tensors are kept
        referenced, so their memory cannot have been handed to another batch in between."""
        e = self.expected
        ids = domain['boundary']
        return e is not None and e[0] is data and e[1] == data._version and e[2] is ids and e[3] == ids._version

    def load(self, data: Tensor, domain: dict) -> None:
        names = list(self.domain)
        srcs = [data] + [domain[k] for k in names]
        if len(srcs) <= 16 and all(t.is_cuda and t.is_contiguous() and t.shape[0] == data.shape[0] and
                                   t.dtype in (torch.float32, torch.int64) for t in srcs):
            ops.copy_blocks_multi(srcs, [self.data] + [self.domain[k] for k in names])      # one launch
            return
        self.data.copy_(data, non_blocking=True)
        for k, v in domain.items():
            self.domain[k].copy_(v, non_blocking=True)

    def run(self, data: Tensor, domain: dict, next_data: Optional[Tensor] = None, next_domain: Optional[dict] = None,
            geo: Optional[list] = None) -> 'StepResult':
        self.load(data, domain)
        if self.cached:
            if geo is None:
                raise _lib.PcfdError('this step graph was captured for batches that bring their cached geometry')
            for cur, new in zip(self.geo, geo):
                for k in cur:
                    cur[k].copy_(new[k], non_blocking=True)
        if self.pipeline:
            if not self._announced(data, domain):
                # not announced by the previous call: geometry of this batch in line, into the static buffers
                fresh = self.ex.geometry(self.data, self.labels, self.domain)
                for cur, new in zip(self.geo, fresh):
                    for k in cur:
                        cur[k].copy_(new[k])
            if next_data is None:
                next_data, next_domain = data, domain
            self.ex.geometry_positions(next_data, self.labels, next_domain, out=self.pos_next)     # one launch
            self.expected = (next_data, next_data._version, next_domain['boundary'], next_domain['boundary']._version)
        self.graph.replay()
        return self.result


class _EagerStep:
    """Stand-in for GraphedStep when capture is not possible."""

    def __init__(self, ex, labels, laplacian):
        self.ex, self.labels, self.laplacian = weakref.proxy(ex), labels, laplacian

    def run(self, data, domain, next_data=None, next_domain=None, geo=None):
        return self.ex.step(data, self.labels, domain, self.laplacian, geo=geo)


class PinnExecutor:
    """Runs forward / training step of one model through the CUDA kernels.  Built lazily by
    `PorousPinnBase` once the model lives on a CUDA device."""

    def __init__(self, model):
        # weak: model -> executor -> model would be a reference cycle, and a dropped model would then keep its CUDA graphs,
        # memory pools and GBs of buffers alive until Python's cyclic collector happens to run (possibly inside a later
        # graph capture, see GraphedStep); with a proxy the executor dies with its model
        self.model = weakref.proxy(model)
        self.device = next(model.parameters()).device
        if self.device.type != 'cuda':
            raise _lib.PcfdError('the porous-cfd hot path only runs on a CUDA sm_100 device; move the model with '
                                 '.to("cuda") first (there is no CPU fallback)')
        _lib.load()
        self.ctx = StepContext(self.device)
        self.params = [p for p in model.parameters()]
        total = sum(p.numel() for p in self.params)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=self.device)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.PcfdError('parameters must be contiguous float32')
            self.ctx.grads[id(p)] = self.flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.plan = model.build_plan()   # dict, see models/*.py
        self._graphs: dict = {}
        self._seen: set = set()

    def rebind_flat_grad(self, buf: Tensor) -> None:
        """Move the flat gradient into caller-provided storage (the data-parallel trainer places it in symmetric memory so
        that peers can read it).  Captured graphs hold the old address and are dropped."""
        if buf.numel() != self.flat_grad.numel() or buf.dtype != torch.float32 or buf.device != self.flat_grad.device:
            raise _lib.PcfdError('rebind_flat_grad: buffer does not match the flat gradient')
        buf.copy_(self.flat_grad)
        self.flat_grad = buf
        off = 0
        for p in self.params:
            self.ctx.grads[id(p)] = self.flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.reset_graphs()

    def reset_graphs(self) -> None:
        """Forget the captured step graphs (their kernels hold loss weights, scaler values and buffer addresses by
        value): the next call of a signature runs eagerly, the one after re-captures."""
        self._graphs.clear()
        self._seen.clear()

    # ---- helpers --------------------------------------------------------------------------
    def _cols(self, labels: dict, name: str):
        keys = list(labels.keys())
        sub = labels[name]
        return [keys.index(s) for s in sub] if sub else [keys.index(name)]

    def _gather(self, data: Tensor, row_ids: Optional[Tensor], n_sel: int, cols, out: Tensor, ldout: int,
                out_rows_per_geom: int, out_row_offset: int = 0, out_col_offset: int = 0):
        b, n_rows, f = data.shape
        ops.gather_cols(data, b, n_rows, f, row_ids, 0, n_sel, cols, out, ldout, out_rows_per_geom, out_row_offset,
                        out_col_offset)

    def _segmax_features(self, ctx, layers, pending_act, zin: Jet, n_seg: int, seg_len: int):
        zs = chain_forward(ctx, layers, zin, seg_len)
        c = layers[-1].n
        out, arg, zsel = ops.segmax_fwd_z(zs[-1].t[0], pending_act, None, n_seg, seg_len, c)
        return out, {'zs': zs, 'arg': arg, 'zsel': zsel, 'n_seg': n_seg, 'seg_len': seg_len, 'c': c, 'act': pending_act}

    def _segmax_features_bwd(self, ctx, layers, sv, gout: Tensor, ldgout: int, need_input_grad=False, side=None):
        return pool_backward(ctx, layers, sv['zs'], gout, ldgout, sv['arg'], sv['zsel'], sv['act'], sv['n_seg'], sv['seg_len'],
                             rows_per_geom=sv['seg_len'], need_input_grad=need_input_grad, side=side)

    # ---- encode: per-geometry constants ------------------------------------------------------
    def _encode(self, data: Tensor, labels: dict, domain: dict, pts_int_ids, pts_bnd_ids, points: Optional[Tensor],
                geo: Optional[list] = None):
        """Returns (cvecs, escale, saved).  `points` (B,N,D) overrides the coordinates taken from
        `data` (used by forward(autograd_points, x)).  `geo`: set-abstraction geometry computed ahead (self.geometry)."""
        plan, ctx = self.plan, self.ctx
        b, n_rows, f = data.shape
        d = plan['dims']
        c_cols = self._cols(labels, 'C')
        saved = {}
        cvecs, escale = {}, None
        fam = plan['family']
        if fam in ('pipn_pp', 'pigano_pp'):
            bnd_ids = domain['boundary']
            nb = bnd_ids.shape[1]
            pos = torch.empty((b, nb, d), dtype=torch.float32, device=data.device)
            self._gather(data, bnd_ids, nb, c_cols, pos, d, nb)
            gcols = []
            for name in plan['geom_feature_order']:
                gcols += self._cols(labels, name)
            x0 = torch.empty((b * nb, ops.round4(len(gcols))), dtype=torch.float32, device=data.device)
            self._gather(data, bnd_ids, nb, gcols, x0, x0.stride(0), nb)
            g, sa_saved = sa_forward(ctx, plan['sa_stack'], x0, x0.stride(0), len(gcols), pos, geo)
            saved['sa'] = sa_saved
            gfeat, gwidth = g, plan['sa_stack'].global_layers[-1].n
        elif fam == 'pigano':
            n_all = n_rows
            gcols = self._cols(labels, 'boundaryId') + self._cols(labels, 'sdf')
            width = len(gcols) + d
            gin = torch.empty((1, b * n_all, ops.round4(width)), dtype=torch.float32, device=data.device)
            self._gather(data, None, n_all, gcols, gin, gin.stride(1), n_all)
            if points is not None:
                ops.gather_cols(points, b, n_all, d, None, 0, n_all, list(range(d)), gin, gin.stride(1), n_all, 0,
                                len(gcols))
            else:
                ni, nbd = pts_int_ids.shape[1], pts_bnd_ids.shape[1]
                self._gather(data, pts_int_ids, ni, c_cols, gin, gin.stride(1), n_all, 0, len(gcols))
                self._gather(data, pts_bnd_ids, nbd, c_cols, gin, gin.stride(1), n_all, ni, len(gcols))
            gfeat, saved['geom'] = self._segmax_features(ctx, plan['geom_layers'], plan['geom_pending_act'],
                                                         Jet(gin, width), b, n_all)
            gwidth = plan['geom_layers'][-1].n
        elif fam == 'pipn':
            gfeat, gwidth, saved['pipn'] = self._encode_pipn(data, labels, pts_int_ids, pts_bnd_ids, points)
        else:
            raise KeyError(fam)

        if fam in ('pigano', 'pigano_pp'):
            # branch network over (sub-domain rows x [C, variable-boundary features]) (models/pi_gano/base.py:60-73)
            vb = plan['variable_boundaries']
            pcols = list(c_cols)
            for name in vb['Features']:
                pcols += self._cols(labels, name)
            counts = [domain[s].shape[1] for s in vb['Subdomains']]
            n_par = sum(counts)
            pin = torch.empty((1, b * n_par, ops.round4(len(pcols))), dtype=torch.float32, device=data.device)
            off = 0
            for s, cnt in zip(vb['Subdomains'], counts):
                self._gather(data, domain[s], cnt, pcols, pin, pin.stride(1), n_par, off)
                off += cnt
            escale, saved['branch'] = self._segmax_features(ctx, plan['branch_layers'], plan['branch_pending_act'],
                                                            Jet(pin, len(pcols)), b, n_par)

        # fold the global feature into the per-geometry constant of the concat layer
        cl = plan['concat_layer']            # ChainLayer over the GLOBAL column block (bias lives here)
        gjet = Jet(gfeat.unsqueeze(0), gwidth)
        cv = ops.jet_linear_fwd(gjet, None, cl.weight, cl.col_lo, cl.k, cl.bias, None, 0, cl.n)
        cvecs['concat'] = cv.t[0]
        saved['gjet'] = gjet
        return cvecs, escale, saved

    def _encode_pipn(self, data, labels, pts_int_ids, pts_bnd_ids, points):
        """Vanilla PIPN global feature (models/modules.py:71-82): local MLP on every point, concat with
        [boundaryId, sdf], global MLP, max over points.  This is the VALUE path; the dependence of the pooled
        feature on the autograd points (max-pool coupling) enters the step through
        coupling.py (`laplacian='reference'`: exact; `'true'`: per-point terms only)."""
        plan, ctx = self.plan, self.ctx
        b, n_rows, f = data.shape
        d = plan['dims']
        c_cols = self._cols(labels, 'C')
        z0 = Jet.empty(1, b * n_rows, d, data.device)
        if points is not None:
            ops.gather_cols(points, b, n_rows, d, None, 0, n_rows, list(range(d)), z0.t, z0.ld, n_rows)
        else:
            ni, nbd = pts_int_ids.shape[1], pts_bnd_ids.shape[1]
            self._gather(data, pts_int_ids, ni, c_cols, z0.t, z0.ld, n_rows, 0)
            self._gather(data, pts_bnd_ids, nbd, c_cols, z0.t, z0.ld, n_rows, ni)
        zs_local = chain_forward(ctx, plan['local_layers'], z0, n_rows)
        lw = plan['local_layers'][-1].n
        gcols = self._cols(labels, 'boundaryId') + self._cols(labels, 'sdf')
        width = lw + len(gcols)
        gin = Jet.empty(1, b * n_rows, width, data.device)
        ops.gather_cols(zs_local[-1].t, 1, b * n_rows, zs_local[-1].ld, None, 0, b * n_rows, list(range(lw)),
                        gin.t, gin.ld, b * n_rows)
        self._gather(data, None, n_rows, gcols, gin.t, gin.ld, n_rows, 0, lw)
        gfeat, sv = self._segmax_features(ctx, plan['global_layers'], plan['global_pending_act'], gin, b, n_rows)
        return gfeat, plan['global_layers'][-1].n, {'zs_local': zs_local, 'global': sv, 'lw': lw}

    def _encode_backward(self, saved: dict, gcvecs: dict, gescale: Optional[Tensor], dw_stream=None, gescale_ready=None):
        """Reverse pass of the encoders on the current stream, their weight-gradient kernels on `dw_stream`.
        `gescale_ready`: event after which `gescale` is complete (the branch part waits for it; the geometry part only
        needs `gcvecs`, which the caller has ordered before this call)."""
        plan, ctx = self.plan, self.ctx
        cur = torch.cuda.current_stream()
        side = dw_stream
        cl = plan['concat_layer']
        gjet = saved['gjet']
        gcv = Jet(gcvecs['concat'].unsqueeze(0), cl.n)
        ops.jet_linear_bwd_dw(gcv, gjet, None, ctx.grad(cl.weight), cl.col_lo, ctx.grad(cl.bias), None, 0, cl.k, cl.n,
                              ctx.need_workspace(ops.dw_workspace_bytes(1, gjet.rows, 0, cl.k, cl.n)))
        gg = ops.jet_linear_bwd_dx(gcv, cl.weight, cl.col_lo, gjet, None, None, 0, cl.k, cl.n)
        ctx.keep.append(gg)
        fam = plan['family']
        if fam in ('pipn_pp', 'pigano_pp'):
            sa_backward(ctx, plan['sa_stack'], saved['sa'], gg.t[0], gg.ld, side=side)
        elif fam == 'pigano':
            self._segmax_features_bwd(ctx, plan['geom_layers'], saved['geom'], gg.t[0], gg.ld, side=side)
        elif fam == 'pipn':
            sv = saved['pipn']
            gin = self._segmax_features_bwd(ctx, plan['global_layers'], sv['global'], gg.t[0], gg.ld,
                                            need_input_grad=True, side=side)
            # gradient of the local features: first lw columns of the concat input, pending activation is
            # applied by the global MLP's first layer (act_cols = lw), so gin[:, :lw] is d/d z_local
            glocal = Jet(gin.t[:, :, :], sv['lw'])
            chain_backward(ctx, plan['local_layers'], sv['zs_local'], glocal, sv['zs_local'][0].rows, side=side)
        if fam in ('pigano', 'pigano_pp'):
            for ev in (gescale_ready or ()):
                cur.wait_event(ev)
            self._segmax_features_bwd(ctx, plan['branch_layers'], saved['branch'], gescale, gescale.stride(0), side=side)
        if side is not None:
            cur.wait_stream(side)

    def _backward_overlapped(self, layers, zs_int, zs_bnd, gy_int: Jet, gy_bnd: Jet, ni: int, nb: int, escale, gescale,
                             gcvecs: dict, saved: dict) -> None:
        """Reverse pass as a dependency graph over five streams (fork / join on the current stream, so it is captured
        into the step graph as parallel branches):

            main    dX chain of the internal points (the large kernels)
            bnd     dX chain of the boundary points (value only, small grids) -- independent of the internal chain
            side    weight gradients of BOTH point chains, layer by layer in one stream order (they accumulate into the
                    same gradient tensors); each waits only for the cotangent it consumes
            enc     reverse pass of the encoders: starts as soon as the concat layer's per-geometry cotangent is complete
                    (both chains' dW of that layer), i.e. while the point chains are still going back through the layers
                    in front of it; the branch (PI-GANO) part waits for the last branch-scaled dX of both chains
            encdw   the encoders' weight gradients

        The encoder's reverse pass is a long chain of small, latency-bound launches (pooling, set-abstraction layers on a
        few thousand rows) and the boundary chain runs one wave per kernel: next to the internal chain they cost nothing,
        one after the other they were ~40 % of the backward (kernel timeline in profiles/r2_timeline_*.md)."""
        ctx = self.ctx
        main = torch.cuda.current_stream()
        s_bnd, s_dw, s_enc, s_encdw = ctx.stream('bnd'), ctx.stream('side'), ctx.stream('enc'), ctx.stream('encdw')
        for st in (s_bnd, s_dw, s_enc):
            st.wait_stream(main)
        ctx.keep += [gy_int, gy_bnd, gescale, gcvecs, zs_int, zs_bnd]
        gz_i, gz_b = gy_int, gy_bnd
        first_escale = min((i for i, L in enumerate(layers) if L.escale), default=None)
        gescale_ready = None
        for i in range(len(layers) - 1, -1, -1):
            L = layers[i]
            tin_i = _tin(ctx, L, escale, 100 + i)
            tin_b = _tin(ctx, L, escale, 200 + i)
            gbias = ctx.grad(L.bias) if L.cvec_key is None else None
            gcvec = gcvecs[L.cvec_key] if L.cvec_key is not None else None
            ev_i, ev_b = torch.cuda.Event(), torch.cuda.Event()
            ev_i.record(main)
            ev_b.record(s_bnd)
            if L.cvec_key is not None:
                # The encoders can go back as soon as the per-geometry cotangent of this layer's constant is known: it is
                # the column sum of the value plane of gz per geometry, so it is formed here by the column-sum kernels
                # alone (gw = None) instead of waiting for this layer's dW in the weight-gradient queue.
                with torch.cuda.stream(s_enc):
                    for ev, gz, zin, rpg in ((ev_i, gz_i, zs_int[i], ni), (ev_b, gz_b, zs_bnd[i], nb)):
                        s_enc.wait_event(ev)
                        g0, z0 = Jet(gz.t[0:1], gz.width), Jet(zin.t[0:1], zin.width)
                        ops.jet_linear_bwd_dw(g0, z0, None, None, 0, None, gcvec, rpg, L.k, L.n,
                                              ctx.need_workspace(ops.dw_workspace_bytes(1, gz.rows, rpg, L.k, L.n)))
                    if first_escale is not None and gescale_ready is None:
                        raise _lib.PcfdError('branch-scaled layers in front of the concat layer are not supported')
                    self._encode_backward(saved, gcvecs, gescale, dw_stream=s_encdw, gescale_ready=gescale_ready)
                gcvec = None
            with torch.cuda.stream(s_dw):
                s_dw.wait_event(ev_i)
                zin = zs_int[i]
                ops.jet_linear_bwd_dw(gz_i, zin, tin_i, ctx.grad(L.weight), L.col_lo, gbias, gcvec, ni, L.k, L.n,
                                      ctx.need_workspace(ops.dw_workspace_bytes(zin.cj, zin.rows, ni, L.k, L.n)))
                s_dw.wait_event(ev_b)
                zin = zs_bnd[i]
                ops.jet_linear_bwd_dw(gz_b, zin, tin_b, ctx.grad(L.weight), L.col_lo, gbias, gcvec, nb, L.k, L.n,
                                      ctx.need_workspace(ops.dw_workspace_bytes(zin.cj, zin.rows, nb, L.k, L.n)))
            if i > 0:
                gz_i = ops.jet_linear_bwd_dx(gz_i, L.weight, L.col_lo, zs_int[i], tin_i, gescale if L.escale else None, ni,
                                             L.k, L.n)
                with torch.cuda.stream(s_bnd):
                    gz_b = ops.jet_linear_bwd_dx(gz_b, L.weight, L.col_lo, zs_bnd[i], tin_b, gescale if L.escale else None,
                                                 nb, L.k, L.n)
                ctx.keep += [gz_i, gz_b]
            if first_escale is not None and i == first_escale:
                # gescale is complete once this layer's dX ran on both chains
                gescale_ready = (torch.cuda.Event(), torch.cuda.Event())
                gescale_ready[0].record(main)
                gescale_ready[1].record(s_bnd)
        for st in (s_bnd, s_dw, s_enc):
            main.wait_stream(st)

    # ---- public entry points -----------------------------------------------------------------
    def forward_values(self, points: Tensor, data: Tensor, labels: dict, domain: dict) -> Tensor:
        """Model.forward(autograd_points, x): predictions (B, N, D+1) at `points` (value pass only)."""
        plan, ctx = self.plan, self.ctx
        ctx.training = self.model.training
        b, n, d = points.shape
        points = points.detach().contiguous().float()
        ops.begin_step()
        cvecs, escale, _ = self._encode(data, labels, domain, None, None, points)
        z0 = Jet.empty(1, b * n, d, data.device)
        ops.gather_cols(points, b, n, d, None, 0, n, list(range(d)), z0.t, z0.ld, n)
        zs = chain_forward(ctx, plan['point_layers'], z0, n, escale, cvecs, salt_base=100)
        ops.end_step()
        return zs[-1].values().reshape(b, n, d + 1)

    def forward_jets(self, points: Tensor, data: Tensor, labels: dict, domain: dict, order: int = 2) -> Jet:
        """Model.forward(autograd_points, x) with its spatial derivatives: the output jet [cj][B*N][ld] at ALL `points`
        (B, N, D): plane 0 the predictions, planes 1..D d/dx_k, planes D+1..2D d2/dx_k2 (order 2).  This is what the
        reference obtains from autograd.grad sweeps over the points (models/model_base.py:11-53); served to
        get_jacobian / get_laplacian / calculate_gradients through model_base._JetForward."""
        plan, ctx = self.plan, self.ctx
        ctx.training = self.model.training
        b, n, d = points.shape
        points = points.detach().contiguous().float()
        cj = 1 + order * d
        ops.begin_step()
        cvecs, escale, _ = self._encode(data, labels, domain, None, None, points)
        z0 = ops.seed_jet(points, None, n, list(range(d)), cj)
        zs = chain_forward(ctx, plan['point_layers'], z0, n, escale, cvecs, salt_base=100)
        ops.end_step()
        return zs[-1]

    def graphed_step(self, data: Tensor, labels: dict, domain: dict, laplacian: str = 'reference',
                     next_batch=None, geo: Optional[list] = None) -> StepResult:
        """`step` replayed from a CUDA graph.  The first call with a given signature runs eagerly (it sizes the
        workspace and the padded weight copies), the second captures, later ones replay; every call is exactly one
        training step (dropout seed, ReLoBRaLo state and gradients advance once)."""
        key = (tuple(data.shape), tuple((k, tuple(v.shape)) for k, v in sorted(domain.items())), tuple(labels),
               laplacian, self.model.training, self.model.enable_data_loss)
        pipeline = bool(getattr(self.model, 'pipeline_geometry', False)) and geo is None
        key = key + (pipeline, geo is not None)
        nd, ndom = (next_batch.data, next_batch.domain) if next_batch is not None else (None, None)
        if nd is not None and (tuple(nd.shape) != tuple(data.shape) or any(tuple(ndom[k].shape) != tuple(v.shape) for k, v in domain.items())):
            nd, ndom = None, None        # a differently shaped next batch (last, smaller one) cannot share the buffers
        g = self._graphs.get(key)
        if g is not None:
            return g.run(data, domain, nd, ndom, geo=geo)
        if key not in self._seen:
            self._seen.add(key)
            return self.step(data, labels, domain, laplacian, geo=geo)
        prev_stream = torch.cuda.current_stream()
        try:
            g = GraphedStep(self, data, labels, domain, laplacian, pipeline, geo=geo)
        except RuntimeError as exc:      # an op that cannot be captured: stay on the per-kernel launches for this signature
            import warnings
            first = exc.__context__ if exc.__context__ is not None else exc       # the error raised INSIDE the capture
            warnings.warn(f'CUDA graph capture of the fused step failed ({type(first).__name__}: {str(first)[:400]}'
                          + (f' -> {str(exc)[:200]}' if first is not exc else '') + '); launching eagerly')
            # a capture that dies half-way leaves torch's current stream on the (ended, invalidated) capture stream and the
            # executor's side streams possibly inside it: go back to the caller's stream and to fresh side streams
            torch.cuda.set_stream(prev_stream)
            self.ctx.streams.clear()
            self.ctx.side_stream = None
            self.ctx.workspaces.clear()
            try:
                torch.cuda.synchronize()
            except RuntimeError:
                pass
            self._seen.discard(key)
            self._graphs[key] = _EagerStep(self, labels, laplacian)
            return self._graphs[key].run(data, domain, geo=geo)
        self._graphs[key] = g
        return g.run(data, domain, nd, ndom, geo=geo)

    def predict_with_residuals(self, data: Tensor, labels: dict, domain: dict, laplacian: str = 'reference'):
        """predict_step(verbose): predictions at all points (B, N, D+1) and the residual map of the internal
        points (B, NI, D+1) = cat([momentum residual, divergence]) (reference models/model_base.py:233-252).
        One jet forward, no reduction, no backward."""
        plan, ctx, model = self.plan, self.ctx, self.model
        ctx.training = model.training
        b, n_rows, f = data.shape
        d = plan['dims']
        int_ids, bnd_ids = domain['internal'], domain['boundary']
        ni, nb = int_ids.shape[1], bnd_ids.shape[1]
        cj = 1 + d if laplacian == 'reference' else 1 + 2 * d
        c_cols = self._cols(labels, 'C')
        ops.begin_step()
        cvecs, escale, _ = self._encode(data, labels, domain, int_ids, bnd_ids, None)
        z0_int = ops.seed_jet(data, int_ids, ni, c_cols, cj)
        z0_bnd = ops.seed_jet(data, bnd_ids, nb, c_cols, 1)
        layers = plan['point_layers']
        y_int = chain_forward(ctx, layers, z0_int, ni, escale, cvecs, salt_base=100)[-1]
        y_bnd = chain_forward(ctx, layers, z0_bnd, nb, escale, cvecs, salt_base=200)[-1]
        ops.end_step()
        fields = ops.residual_fields(data, int_ids, y_int, model.residual_params(labels, laplacian))
        pred = torch.cat([y_int.values().reshape(b, ni, d + 1), y_bnd.values().reshape(b, nb, d + 1)], dim=1)
        return pred, fields

    def uses_geometry(self) -> bool:
        """True for the models with a set-abstraction encoder (FPS / ball query in front of the step)."""
        return self.plan['family'] in ('pipn_pp', 'pigano_pp')

    def geometry_positions(self, data: Tensor, labels: dict, domain: dict, out: Optional[Tensor] = None) -> Tensor:
        """(B, n_boundary, D) coordinates of the points the set-abstraction encoder samples (one gather launch)."""
        b = data.shape[0]
        d = self.plan['dims']
        bnd_ids = domain['boundary']
        nb = bnd_ids.shape[1]
        pos = out if out is not None else torch.empty((b, nb, d), dtype=torch.float32, device=data.device)
        self._gather(data, bnd_ids, nb, self._cols(labels, 'C'), pos, d, nb)
        return pos

    def geometry(self, data: Tensor, labels: dict, domain: dict, pos: Optional[Tensor] = None) -> Optional[list]:
        """The weight-independent head of the encoder for a batch: FPS centroids, neighbourhood slots and centroid
        positions of every set-abstraction level (sa_geometry).  Pass the result to `step(..., geo=...)`."""
        if not self.uses_geometry():
            return None
        if pos is None:
            pos = self.geometry_positions(data, labels, domain)
        return sa_geometry(self.plan['sa_stack'], pos)

    def step(self, data: Tensor, labels: dict, domain: dict, laplacian: str = 'reference',
             keep_outputs: bool = False, geo: Optional[list] = None, accumulate: bool = False) -> StepResult:
        """One fused training step: fills the flat gradient buffer and returns the loss vector.  `accumulate`: add to
        the gradient buffer instead of overwriting it (micro-batches of one optimizer step)."""
        plan, ctx = self.plan, self.ctx
        model = self.model
        ctx.training = model.training
        if data.dtype != torch.float32 or not data.is_contiguous():
            raise _lib.PcfdError('FoamData.data must be contiguous float32')
        b, n_rows, f = data.shape
        d = plan['dims']
        int_ids, bnd_ids = domain['internal'], domain['boundary']
        ni, nb = int_ids.shape[1], bnd_ids.shape[1]
        obs_ids = domain.get('obs') if model.enable_data_loss else None
        if obs_ids is not None and obs_ids.shape[1] == 0:
            obs_ids = None
        cj = 1 + d if laplacian == 'reference' else 1 + 2 * d
        c_cols = self._cols(labels, 'C')

        ops.begin_step()
        ctx.keep.clear()
        if not accumulate:
            ops.zero_(self.flat_grad)
        if ctx.training:
            ops.advance_seed(ctx.seed_dev)

        layers = plan['point_layers']
        # The point chains up to the concat layer depend on the coordinates only, the encoder (FPS, ball query, set
        # abstraction: a long chain of small, latency-bound launches) on nothing they produce: run the former on a
        # side stream while the encoder occupies the main one (fork / join, also inside a captured graph).
        n_pre = next((i for i, L in enumerate(layers) if L.cvec_key is not None or L.escale), len(layers))
        main = torch.cuda.current_stream()
        ctx.side_stream = ctx.stream('side')
        side = ctx.side_stream if ctx.overlap else main
        side.wait_stream(main)
        with torch.cuda.stream(side):
            z0_int = ops.seed_jet(data, int_ids, ni, c_cols, cj)
            z0_bnd = ops.seed_jet(data, bnd_ids, nb, c_cols, 1)
            zs_int = chain_forward(ctx, layers[:n_pre], z0_int, ni, None, None, salt_base=100)
            zs_bnd = chain_forward(ctx, layers[:n_pre], z0_bnd, nb, None, None, salt_base=200)
        cvecs, escale, saved = self._encode(data, labels, domain, int_ids, bnd_ids, None, geo)
        main.wait_stream(side)
        # the boundary chain (value only, small grids) fills the gaps of the internal chain from the side stream
        side.wait_stream(main)
        with torch.cuda.stream(side):
            zs_bnd += chain_forward(ctx, layers[n_pre:], zs_bnd[-1], nb, escale, cvecs, salt_base=200 + n_pre)[1:]
        zs_int += chain_forward(ctx, layers[n_pre:], zs_int[-1], ni, escale, cvecs, salt_base=100 + n_pre)[1:]
        main.wait_stream(side)

        coup = None
        if plan['family'] == 'pipn' and laplacian == 'reference' and getattr(model, 'coupling', True):
            # vanilla PIPN: max-pool cross-point terms of the reference's summed-output Jacobian (coupling.py)
            from . import coupling
            coup = coupling.forward(self, data, labels, int_ids, zs_int, saved, cj,
                                    list(model.residual_params(labels, laplacian).c_std))

        prm = model.residual_params(labels, laplacian)
        ctx.need_workspace(ops.residual_workspace_bytes(b, ni, nb, obs_ids.shape[1] if obs_ids is not None else 0))
        scaler = getattr(model, 'loss_scaler', None)
        weights_dev = None
        if scaler is not None and getattr(scaler, 'dynamic', False) and ctx.training:
            # adaptive weights (ReLoBRaLo): unscaled terms first, buffer / weight update on the device, then the weighted
            # pass with the device-resident weights (models/losses.py:93-124)
            n_terms = 2 * d + 2 + ((d + 1) if obs_ids is not None else 0)
            _, _, out0 = ops.residual_loss(data, int_ids, bnd_ids, obs_ids, zs_int[-1], zs_bnd[-1], prm, ctx.workspace)
            step_dev, weights_dev = scaler.device_state(data.device)
            ops.relobralo_update(out0, n_terms, scaler.init_losses, scaler.prev_losses, scaler.lambda_ema, step_dev,
                                 scaler.batch_size, scaler.alpha, scaler.beta, scaler.tau, scaler.eps, scaler.seed,
                                 weights_dev)
        gy_int, gy_bnd, out = ops.residual_loss(data, int_ids, bnd_ids, obs_ids, zs_int[-1], zs_bnd[-1], prm,
                                                ctx.workspace, weights_dev,
                                                coup.visc_extra if coup is not None else None,
                                                coup.gvisc if coup is not None else None)

        gcvecs = {'concat': torch.empty_like(cvecs['concat'])}
        ops.zero_(gcvecs['concat'])
        gescale = None
        if escale is not None:
            gescale = torch.empty_like(escale)
            ops.zero_(gescale)
        if coup is None and ctx.overlap:
            self._backward_overlapped(layers, zs_int, zs_bnd, gy_int, gy_bnd, ni, nb, escale, gescale, gcvecs, saved)
        else:
            # one stream (per-kernel timing), or the vanilla-PIPN coupling pass whose reverse sweeps interleave with these
            chain_backward(ctx, layers, zs_int, gy_int, ni, escale, gescale, gcvecs, salt_base=100)
            chain_backward(ctx, layers, zs_bnd, gy_bnd, nb, escale, gescale, gcvecs, salt_base=200)
            if coup is not None:
                from . import coupling
                coupling.backward(self, coup, data, int_ids, zs_int, gy_int, saved, gcvecs)
            self._encode_backward(saved, gcvecs, gescale, dw_stream=ctx.stream('side') if ctx.overlap else None)
        ops.end_step()

        n_terms = 2 * d + 2 + ((d + 1) if obs_ids is not None else 0)
        res = StepResult(out, n_terms)
        if keep_outputs:
            res.y_int, res.y_bnd = zs_int[-1], zs_bnd[-1]
        return res
