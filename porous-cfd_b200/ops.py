"""Tensor-level wrappers of the C ABI (include/pcfd.h).  Each function takes CUDA tensors, passes
raw device pointers and the current torch stream to libpcfd_sm100.so and raises on a non-zero
status.  No arithmetic happens in Python or in torch kernels here."""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional

import torch
from torch import Tensor

from . import _lib
from ._lib import ACT_CODES, InTrans, ResidualParams, check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class KernelProfile:
    """Optional per-kernel-family timing with CUDA events on the launching stream (bench.py's
    roofline leg).  `work` is the algorithmic FLOP (GEMM families) or byte count of the launch."""

    def __init__(self):
        self.records = []   # (name, start_event, end_event, work, bytes)

    def summary(self) -> dict:
        torch.cuda.synchronize()
        out = {}
        for name, a, b, work, nbytes in self.records:
            d = out.setdefault(name, {'launches': 0, 'ms': 0.0, 'work': 0.0, 'bytes': 0.0})
            d['launches'] += 1
            d['ms'] += a.elapsed_time(b)
            d['work'] += work
            d['bytes'] += nbytes
        return out


PROFILE: Optional[KernelProfile] = None


class _timed:
    __slots__ = ('name', 'work', 'nbytes', 'a')

    def __init__(self, name: str, work: float = 0.0, nbytes: float = 0.0):
        """work = algorithmic FLOP (jet GEMMs) or bytes (streaming kernels); nbytes = algorithmic HBM bytes of a GEMM"""
        self.name, self.work, self.nbytes = name, work, nbytes

    def __enter__(self):
        if PROFILE is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            PROFILE.records.append((self.name, self.a, b, self.work, self.nbytes))
        return False


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32(t: Tensor, name: str) -> Tensor:
    if t.dtype != torch.float32 or not t.is_cuda:
        raise _lib.PcfdError(f'{name}: expected a CUDA float32 tensor, got {t.dtype} on {t.device}')
    return t


def round4(n: int) -> int:
    return (n + 3) // 4 * 4


def make_intrans(act=None, act_cols: int = 0, escale: Optional[Tensor] = None, drop_p: float = 0.0,
                 seed_dev: Optional[Tensor] = None, salt: int = 0) -> InTrans:
    t = InTrans()
    t.act = ACT_CODES[act] if not isinstance(act, int) else act
    t.act_cols = act_cols
    t.escale = _ptr(escale)
    t.ldescale = escale.stride(0) if escale is not None else 0
    t.drop_p = float(drop_p)
    t.seed_dev = _ptr(seed_dev) if drop_p > 0 else None
    t.salt = salt
    return t


class Jet:
    """A jet tensor [cj][rows][ld] (ld = width rounded up to 4 floats for vector stores)."""
    __slots__ = ('t', 'cj', 'rows', 'width')

    def __init__(self, t: Tensor, width: int):
        self.t, self.cj, self.rows, self.width = t, t.shape[0], t.shape[1], width

    @staticmethod
    def empty(cj: int, rows: int, width: int, device) -> 'Jet':
        return Jet(torch.empty((cj, rows, round4(width)), dtype=torch.float32, device=device), width)

    @property
    def ld(self) -> int:
        return self.t.stride(1)

    @property
    def plane_stride(self) -> int:
        return self.t.stride(0)

    def values(self) -> Tensor:
        return self.t[0, :, :self.width]


# The TMA-fed engine needs 16-byte aligned weight rows.  Reference parameter shapes such as [64, 10] or
# [128, 131] (set-abstraction MLPs, models/modules.py:506-512) are not, so the step keeps a zero-padded copy of
# those few weight blocks, refreshed once per step by one device copy (inside the captured graph).
_WPAD: dict = {}
_WPAD_FRESH: set = set()
_IN_STEP = False


def begin_step() -> None:
    """Weights do not change until end_step(): padded copies made from here on are reused."""
    global _IN_STEP
    _WPAD_FRESH.clear()
    _IN_STEP = True


def end_step() -> None:
    global _IN_STEP
    _WPAD_FRESH.clear()
    _IN_STEP = False


def _weight_block(w: Tensor, col_lo: int, k: int, n: int):
    if k < 8 or (w.stride(0) % 4 == 0 and col_lo % 4 == 0 and w.data_ptr() % 16 == 0):
        return w.data_ptr() + 4 * col_lo, w.stride(0)
    key = (w.data_ptr(), col_lo, k, n, _stream())     # per stream: the copy is ordered with its consumers
    buf = _WPAD.get(key)
    if buf is None:
        buf = torch.zeros((n, round4(k)), dtype=torch.float32, device=w.device)
        _WPAD[key] = buf
    if key not in _WPAD_FRESH:
        buf[:, :k].copy_(w.detach()[:, col_lo:col_lo + k])
        if _IN_STEP:
            _WPAD_FRESH.add(key)
        _lib.launches += 1
    return buf.data_ptr(), buf.stride(0)


# PCFD_ENGINE=0 (or ops.FORCE_FFMA = True) sends every jet layer to the generic fp32 CUDA-core engine under its own entry
# points (pcfd_ffma_*): the independent implementation tests and scripts/bench_layers.py compare the tensor-core kernels
# with.  The library itself has no engine switch.
FORCE_FFMA = os.environ.get('PCFD_ENGINE', '2') == '0'
# With AUDIT on, every jet layer call first asks pcfd_jet_linear_engine which kernel family will run it; layers of
# tensor-core size that would fall to the generic FFMA engine are recorded in FALLBACKS (bench.py asserts it stays empty).
AUDIT = False
FALLBACKS: list = []


def _audit(pass_id: int, a: 'Jet', w_ptr: int, ldw: int, b: 'Jet', tin, has_gescale: bool, k: int, n: int) -> None:
    if not AUDIT or FORCE_FFMA:
        return
    eng = _lib.load().pcfd_jet_linear_engine(pass_id, a.t.data_ptr(), a.plane_stride, a.ld, w_ptr, ldw, b.t.data_ptr(),
                                             b.plane_stride, b.ld, C.byref(tin) if tin is not None else None,
                                             1 if has_gescale else 0, a.cj, a.rows, k, n)
    if eng == _lib.ENGINE_FFMA and a.rows >= 512 and k >= 8 and n >= 16:      # 512 rows: the smallest dW the tensor-core kernel takes
        FALLBACKS.append((('fwd', 'dx', 'dw')[pass_id], a.cj, a.rows, k, n))


def jet_linear_fwd(zin: Jet, tin: Optional[InTrans], w: Tensor, col_lo: int, k: int, bias: Optional[Tensor],
                   cvec: Optional[Tensor], rows_per_geom: int, n: int, out: Optional[Jet] = None) -> Jet:
    lib = _lib.load()
    if out is None:
        out = Jet.empty(zin.cj, zin.rows, n, zin.t.device)
    _lib.launches += 1
    wptr, ldw = _weight_block(w, col_lo, k, n)
    _audit(0, zin, wptr, ldw, out, tin, False, k, n)
    fn = lib.pcfd_ffma_jet_linear_fwd if FORCE_FFMA else lib.pcfd_jet_linear_fwd
    with _timed(f'jet_fwd_cj{zin.cj}', 2.0 * zin.cj * zin.rows * k * n, 4.0 * zin.cj * zin.rows * (k + n) + 4.0 * k * n):
      check(fn(zin.t.data_ptr(), zin.plane_stride, zin.ld, C.byref(tin) if tin is not None else None,
                                  wptr, ldw, _ptr(bias), _ptr(cvec),
                                  cvec.stride(0) if cvec is not None else 0,
                                  out.t.data_ptr(), out.plane_stride, out.ld, zin.cj, zin.rows, rows_per_geom, k, n,
                                  _stream()), 'pcfd_jet_linear_fwd')
    return out


def jet_linear_bwd_dx(gzout: Jet, w: Tensor, col_lo: int, zin: Jet, tin: Optional[InTrans],
                      gescale: Optional[Tensor], rows_per_geom: int, k: int, n: int) -> Jet:
    lib = _lib.load()
    gzin = Jet.empty(zin.cj, zin.rows, k, zin.t.device)
    _lib.launches += 1
    wptr, ldw = _weight_block(w, col_lo, k, n)
    _audit(1, gzout, wptr, ldw, zin, tin, gescale is not None, k, n)
    fn = lib.pcfd_ffma_jet_linear_bwd_dx if FORCE_FFMA else lib.pcfd_jet_linear_bwd_dx
    with _timed(f'jet_dx_cj{zin.cj}', 2.0 * zin.cj * zin.rows * k * n,
                4.0 * zin.cj * zin.rows * (n + (2 * k if tin is not None else k)) + 4.0 * k * n):
      check(fn(gzout.t.data_ptr(), gzout.plane_stride, gzout.ld, wptr,
                                     ldw, zin.t.data_ptr(), zin.plane_stride, zin.ld,
                                     C.byref(tin) if tin is not None else None,
                                     gzin.t.data_ptr(), gzin.plane_stride, gzin.ld, _ptr(gescale),
                                     gescale.stride(0) if gescale is not None else 0,
                                     zin.cj, zin.rows, rows_per_geom, k, n, _stream()), 'pcfd_jet_linear_bwd_dx')
    return gzin


def dw_workspace_bytes(cj: int, rows: int, rows_per_geom: int, k: int, n: int) -> int:
    return int(_lib.load().pcfd_jet_linear_bwd_dw_workspace_bytes(cj, rows, rows_per_geom, k, n))


def jet_linear_bwd_dw(gzout: Jet, zin: Jet, tin: Optional[InTrans], gw: Optional[Tensor], col_lo: int,
                      gbias: Optional[Tensor], gcvec: Optional[Tensor], rows_per_geom: int, k: int, n: int,
                      workspace: Tensor) -> None:
    lib = _lib.load()
    _lib.launches += 2 + (1 if (gbias is not None or gcvec is not None) else 0)
    if gw is not None:
        _audit(2, gzout, 0, 0, zin, tin, False, k, n)
    fn = lib.pcfd_ffma_jet_linear_bwd_dw if FORCE_FFMA else lib.pcfd_jet_linear_bwd_dw
    with _timed(f'jet_dw_cj{zin.cj}', 2.0 * zin.cj * zin.rows * k * n, 4.0 * zin.cj * zin.rows * (k + n) + 4.0 * k * n):
      check(fn(gzout.t.data_ptr(), gzout.plane_stride, gzout.ld, zin.t.data_ptr(),
                                     zin.plane_stride, zin.ld, C.byref(tin) if tin is not None else None,
                                     (gw.data_ptr() + 4 * col_lo) if gw is not None else None,
                                     gw.stride(0) if gw is not None else 0, _ptr(gbias), _ptr(gcvec),
                                     gcvec.stride(0) if gcvec is not None else 0, zin.cj, zin.rows, rows_per_geom, k, n,
                                     workspace.data_ptr(), workspace.numel() * workspace.element_size(), _stream()),
          'pcfd_jet_linear_bwd_dw')


def segmax_fwd(z: Tensor, act, slots: Optional[Tensor], n_seg: int, seg_len: int, c: int):
    """z [(n_seg*seg_len), ld] pre-activations -> (out [n_seg, c], arg [n_seg, c] int32)."""
    lib = _lib.load()
    out = torch.empty((n_seg, round4(c)), dtype=torch.float32, device=z.device)
    arg = torch.empty((n_seg, c), dtype=torch.int32, device=z.device)
    _lib.launches += 1
    with _timed('segmax_fwd', 4.0 * n_seg * seg_len * c):
      check(lib.pcfd_segmax_fwd(z.data_ptr(), z.stride(0), ACT_CODES[act], _ptr(slots), n_seg, seg_len, c,
                              out.data_ptr(), out.stride(0), arg.data_ptr(), _stream()), 'pcfd_segmax_fwd')
    return out, arg


def segmax_fwd_z(z: Tensor, act, slots: Optional[Tensor], n_seg: int, seg_len: int, c: int):
    """segmax_fwd that also returns zsel [n_seg, c]: the pre-activation of each selected row (for pool_layer_bwd)."""
    lib = _lib.load()
    out = torch.empty((n_seg, round4(c)), dtype=torch.float32, device=z.device)
    arg = torch.empty((n_seg, c), dtype=torch.int32, device=z.device)
    zsel = torch.empty((n_seg, c), dtype=torch.float32, device=z.device)
    _lib.launches += 1
    with _timed('segmax_fwd', 4.0 * n_seg * seg_len * c):
      check(lib.pcfd_segmax_fwd_z(z.data_ptr(), z.stride(0), ACT_CODES[act], _ptr(slots), n_seg, seg_len, c,
                                  out.data_ptr(), out.stride(0), arg.data_ptr(), zsel.data_ptr(), zsel.stride(0), _stream()),
            'pcfd_segmax_fwd_z')
    return out, arg, zsel


def segmax_bwd(gout: Tensor, ldgout: int, arg: Tensor, z: Tensor, act, n_seg: int, seg_len: int, c: int) -> Tensor:
    lib = _lib.load()
    gz = torch.empty((1, n_seg * seg_len, z.stride(0)), dtype=torch.float32, device=z.device)
    _lib.launches += 1
    with _timed('segmax_bwd', 8.0 * n_seg * seg_len * c):
      check(lib.pcfd_segmax_bwd(gout.data_ptr(), ldgout, arg.data_ptr(), z.data_ptr(), z.stride(0), ACT_CODES[act],
                              n_seg, seg_len, c, gz.data_ptr(), gz.stride(1), _stream()), 'pcfd_segmax_bwd')
    return gz


def pool_layer_bwd_supported(n_seg: int, seg_len: int, k: int, c: int, tin: Optional[InTrans], ldzin: int) -> bool:
    return bool(_lib.load().pcfd_pool_layer_bwd_supported(n_seg, seg_len, k, c, C.byref(tin) if tin is not None else None,
                                                          ldzin))


def pool_layer_bwd_workspace_bytes(n_seg: int, seg_len: int, k: int, c: int) -> int:
    return int(_lib.load().pcfd_pool_layer_bwd_workspace_bytes(n_seg, seg_len, k, c))


def pool_layer_bwd(gout: Tensor, ldgout: int, arg: Tensor, zsel: Tensor, act_pool, n_seg: int, seg_len: int, c: int,
                   zin: Jet, tin: Optional[InTrans], k: int, w: Tensor, gw: Optional[Tensor], gbias: Optional[Tensor],
                   need_gzin: bool, workspace: Optional[Tensor]) -> Optional[Jet]:
    """Sparse backward of (last MLP layer -> max pool): accumulates gw / gbias, returns the gradient of the layer's
    input pre-activations (Jet, cj = 1) when `need_gzin`.  zsel [n_seg, c] = the layer's output at the selected rows
    (segmax_fwd_z), zin its inputs; w is the [c][k] weight (row stride w.stride(0))."""
    lib = _lib.load()
    gzin = Jet.empty(1, zin.rows, k, zin.t.device) if need_gzin else None
    want_w = gw is not None or gbias is not None
    _lib.launches += (2 if want_w else 0) + (1 if need_gzin else 0)
    with _timed('pool_layer_bwd', 2.0 * n_seg * c * k * ((1 if want_w else 0) + (1 if need_gzin else 0)),
                4.0 * n_seg * seg_len * k * ((1 if want_w else 0) + (2 if need_gzin else 0)) + 16.0 * n_seg * c):
      check(lib.pcfd_pool_layer_bwd(gout.data_ptr(), ldgout, arg.data_ptr(), zsel.data_ptr(), zsel.stride(0), ACT_CODES[act_pool],
                                    n_seg, seg_len, c, zin.t.data_ptr(), zin.ld, C.byref(tin) if tin is not None else None,
                                    k, w.data_ptr(), w.stride(0), _ptr(gw), gw.stride(0) if gw is not None else 0,
                                    _ptr(gbias), gzin.t.data_ptr() if gzin is not None else None,
                                    gzin.ld if gzin is not None else 0, _ptr(workspace),
                                    workspace.numel() * workspace.element_size() if workspace is not None else 0,
                                    _stream()), 'pcfd_pool_layer_bwd')
    return gzin


def pool_compact(gout: Tensor, ldgout: int, arg: Tensor, zsel: Tensor, act_pool, n_seg: int, c: int):
    """-> (ids int64 [n_seg, c] rows inside each segment, gzc Jet [1, n_seg*c, ld] cotangent of the compacted rows)."""
    lib = _lib.load()
    ids = torch.empty((n_seg, c), dtype=torch.int64, device=zsel.device)
    gzc = Jet.empty(1, n_seg * c, c, zsel.device)
    _lib.launches += 1
    check(lib.pcfd_pool_compact(gout.data_ptr(), ldgout, arg.data_ptr(), zsel.data_ptr(), zsel.stride(0), ACT_CODES[act_pool],
                                n_seg, c, ids.data_ptr(), gzc.t.data_ptr(), gzc.ld, _stream()), 'pcfd_pool_compact')
    return ids, gzc


def fps(pos: Tensor, ratio: float) -> Tensor:
    """pos (B, n, D) -> int64 (B, m) indices into the flattened (B*n) point array, m = ceil(ratio*n)."""
    lib = _lib.load()
    pos = _f32(pos, 'pos').contiguous()
    b, n, d = pos.shape
    m = int(math.ceil(ratio * n))
    idx = torch.empty((b, m), dtype=torch.int64, device=pos.device)
    _lib.launches += 1
    need = int(lib.pcfd_fps_workspace_bytes(b, n, d))
    ws = torch.empty(need, dtype=torch.uint8, device=pos.device) if need else None
    with _timed('fps', 4.0 * b * n * d + 8.0 * b * m):
      check(lib.pcfd_fps_ws(pos.data_ptr(), b, n, d, m, idx.data_ptr(), _ptr(ws), need, _stream()), 'pcfd_fps_ws')
    return idx


def ball_query(pos: Tensor, centroid_idx: Tensor, r: float, k: int):
    """-> (nbr int32 (B*m, k) flattened point indices or -1, count int32 (B*m,))"""
    lib = _lib.load()
    b, n, d = pos.shape
    m = centroid_idx.shape[1]
    nbr = torch.empty((b * m, k), dtype=torch.int32, device=pos.device)
    count = torch.empty((b * m,), dtype=torch.int32, device=pos.device)
    _lib.launches += 1
    with _timed('ball_query', 4.0 * b * n * d + 4.0 * b * m * d + 4.0 * b * m * k):
      check(lib.pcfd_ball_query(pos.data_ptr(), centroid_idx.data_ptr(), b, n, d, m, float(r), k, nbr.data_ptr(),
                              count.data_ptr(), _stream()), 'pcfd_ball_query')
    return nbr, count


def sa_edges(nbr: Tensor, n_points_total: int) -> Tensor:
    lib = _lib.load()
    m_total, k = nbr.shape
    slots = torch.empty((m_total, k + 1), dtype=torch.int32, device=nbr.device)
    _lib.launches += 1
    check(lib.pcfd_sa_edges(nbr.data_ptr(), m_total, k, n_points_total, slots.data_ptr(), _stream()), 'pcfd_sa_edges')
    return slots


def sa_cached_geometry(idx_local: Tensor, nbr_local: Tensor, n: int, idx_out: Optional[Tensor] = None,
                       slots_out: Optional[Tensor] = None):
    """idx_local (B, m) int64, nbr_local (B, m, k) int32 (indices local to each geometry of n points) -> (idx (B, m) int64
    flattened over the batch, slots (B*m, k+1) int32) -- what fps + ball_query + sa_edges produce for the same batch."""
    lib = _lib.load()
    b, m = idx_local.shape
    k = nbr_local.shape[2]
    idx = idx_out if idx_out is not None else torch.empty((b, m), dtype=torch.int64, device=idx_local.device)
    slots = slots_out if slots_out is not None else torch.empty((b * m, k + 1), dtype=torch.int32, device=idx_local.device)
    _lib.launches += 1
    check(lib.pcfd_sa_cached_geometry(idx_local.data_ptr(), nbr_local.data_ptr(), b, m, k, n, idx.data_ptr(), slots.data_ptr(),
                                      _stream()), 'pcfd_sa_cached_geometry')
    return idx, slots


def sa_gather(x: Optional[Tensor], ldx: int, f_in: int, pos: Tensor, centroid_idx: Tensor, slots: Tensor,
              r: float) -> Tensor:
    lib = _lib.load()
    m_total, kp = slots.shape
    dims = pos.shape[-1]
    width = f_in + dims
    ein = torch.empty((1, m_total * kp, round4(width)), dtype=torch.float32, device=pos.device)
    _lib.launches += 1
    with _timed('sa_gather', 8.0 * m_total * kp * width):
      check(lib.pcfd_sa_gather(_ptr(x), ldx, f_in, pos.data_ptr(), dims, centroid_idx.data_ptr(), slots.data_ptr(),
                             m_total, kp, float(r), ein.data_ptr(), ein.stride(1), _stream()), 'pcfd_sa_gather')
    return ein


def sa_scatter_bwd(gein: Tensor, ldgein: int, slots: Tensor, f_in: int, gx: Tensor, ldgx: int) -> None:
    lib = _lib.load()
    m_total, kp = slots.shape
    _lib.launches += 1
    check(lib.pcfd_sa_scatter_bwd(gein.data_ptr(), ldgein, slots.data_ptr(), m_total, kp, f_in, gx.data_ptr(), ldgx,
                                  _stream()), 'pcfd_sa_scatter_bwd')


def gather_cols(data: Tensor, n_geom: int, n_rows: int, f: int, row_ids: Optional[Tensor], first_row: int, n_sel: int,
                cols, out: Tensor, ldout: int, out_rows_per_geom: int, out_row_offset: int = 0,
                out_col_offset: int = 0) -> None:
    lib = _lib.load()
    cols = list(cols)
    arr = None if cols == list(range(len(cols))) else (C.c_int32 * len(cols))(*cols)
    _lib.launches += 1
    check(lib.pcfd_gather_cols(data.data_ptr(), n_geom, n_rows, f, _ptr(row_ids), first_row, n_sel, arr, len(cols),
                               out.data_ptr(), ldout, out_rows_per_geom, out_row_offset, out_col_offset, _stream()),
          'pcfd_gather_cols')


def seed_jet(data: Tensor, row_ids: Optional[Tensor], n_sel: int, coord_cols, cj: int) -> Jet:
    lib = _lib.load()
    b, n_rows, f = data.shape
    dims = len(coord_cols)
    z = Jet.empty(cj, b * n_sel, dims, data.device)
    arr = (C.c_int32 * dims)(*coord_cols)
    _lib.launches += 1
    check(lib.pcfd_seed_jet(data.data_ptr(), b, n_rows, f, _ptr(row_ids), n_sel, arr, dims, cj, z.t.data_ptr(),
                            z.plane_stride, z.ld, _stream()), 'pcfd_seed_jet')
    return z


def residual_workspace_bytes(n_geom: int, ni: int, nb: int, no: int) -> int:
    return int(_lib.load().pcfd_residual_workspace_bytes(n_geom, ni, nb, no))


def residual_loss(data: Tensor, internal_ids: Tensor, boundary_ids: Tensor, obs_ids: Optional[Tensor],
                  y_int: Jet, y_bnd: Jet, prm: ResidualParams, workspace: Tensor, weights_dev: Optional[Tensor] = None,
                  visc_extra: Optional[Tensor] = None, gvisc: Optional[Tensor] = None):
    """-> (gy_int Jet, gy_bnd Jet, out float32[48]).  `weights_dev`: device-resident loss weights (adaptive
    scaler) instead of prm.weights."""
    lib = _lib.load()
    b, n_rows, f = data.shape
    ni, nb = internal_ids.shape[1], boundary_ids.shape[1]
    no = obs_ids.shape[1] if obs_ids is not None else 0
    # one block [gy_bnd | gy_int]: the value planes the fused kernel accumulates into are adjacent -> one memset node
    n_bnd, n_int = y_bnd.t.numel(), y_int.t.numel()
    block = torch.empty(n_bnd + n_int, dtype=torch.float32, device=data.device)
    gy_bnd = Jet(block[:n_bnd].view(y_bnd.t.shape), y_bnd.width)
    gy_int = Jet(block[n_bnd:].view(y_int.t.shape), y_int.width)
    out = torch.empty(_lib.LOSS_OUT_FLOATS, dtype=torch.float32, device=data.device)
    with _timed('residual_loss', 8.0 * b * ni * y_int.cj * y_int.ld + 4.0 * b * ni * f + 8.0 * b * nb * y_int.ld):
      if FUSED_RESIDUAL and y_int.t.is_contiguous() and y_bnd.t.is_contiguous():
        _lib.launches += 1
        check(lib.pcfd_residual_step(data.data_ptr(), b, n_rows, f, internal_ids.data_ptr(), ni, boundary_ids.data_ptr(),
                                     nb, _ptr(obs_ids) if no > 0 else None, no, y_int.t.data_ptr(), y_int.plane_stride,
                                     y_bnd.t.data_ptr(), y_int.ld, C.byref(prm), _ptr(weights_dev), _ptr(visc_extra),
                                     _ptr(gvisc), gy_int.t.data_ptr(), gy_bnd.t.data_ptr(), out.data_ptr(),
                                     _residual_ticket(data.device).data_ptr(), workspace.data_ptr(),
                                     workspace.numel() * workspace.element_size(), _stream()), 'pcfd_residual_step')
      else:
        _lib.launches += 4
        check(lib.pcfd_residual_loss_w(data.data_ptr(), b, n_rows, f, internal_ids.data_ptr(), ni, boundary_ids.data_ptr(),
                                   nb, _ptr(obs_ids) if no > 0 else None, no, y_int.t.data_ptr(), y_int.plane_stride,
                                   y_bnd.t.data_ptr(), y_int.ld, C.byref(prm), _ptr(weights_dev), _ptr(visc_extra),
                                   _ptr(gvisc), gy_int.t.data_ptr(),
                                   gy_bnd.t.data_ptr(), out.data_ptr(), workspace.data_ptr(),
                                   workspace.numel() * workspace.element_size(), _stream()), 'pcfd_residual_loss_w')
    return gy_int, gy_bnd, out


# PCFD_FUSED_RESIDUAL=0: the four-launch form of the residual stage (pcfd_residual_loss_w), kept as the independent
# implementation the fused kernel is tested against
FUSED_RESIDUAL = os.environ.get('PCFD_FUSED_RESIDUAL', '1') != '0'
_TICKETS: dict = {}


def _residual_ticket(device) -> Tensor:
    """The last-block ticket of pcfd_residual_step: one zero-initialised int32 per (device, stream); the kernel hands it
    back at zero."""
    key = (str(device), _stream())
    t = _TICKETS.get(key)
    if t is None:
        t = _TICKETS[key] = torch.zeros(1, dtype=torch.int32, device=device)
    return t


def residual_fields(data: Tensor, internal_ids: Tensor, y_int: Jet, prm: ResidualParams) -> Tensor:
    """-> (B, NI, D+1) = cat([momentum residual, divergence]) at the internal points (predict_step, verbose)."""
    lib = _lib.load()
    b, n_rows, f = data.shape
    ni = internal_ids.shape[1]
    d = prm.dims
    out = torch.empty((b, ni, d + 1), dtype=torch.float32, device=data.device)
    _lib.launches += 1
    check(lib.pcfd_residual_fields(data.data_ptr(), b, n_rows, f, internal_ids.data_ptr(), ni, y_int.t.data_ptr(),
                                   y_int.plane_stride, y_int.ld, C.byref(prm), out.data_ptr(), _stream()),
          'pcfd_residual_fields')
    return out


def residual_eval(prm: ResidualParams, jac: Tensor, u: Optional[Tensor] = None, lap: Optional[Tensor] = None,
                  p_grad: Optional[Tensor] = None, zone: Optional[Tensor] = None, dcoef: Optional[Tensor] = None,
                  fcoef: Optional[Tensor] = None, want_momentum: bool = True, want_div: bool = False):
    """Loss-module `func` on explicit tensors (pcfd_residual_eval): jac (..., D, D) [, u (..., D), lap (..., D, D),
    p_grad (..., D), zone (..., 1), dcoef / fcoef (..., D)] -> (momentum (..., D) or None, div (...) or None)."""
    lib = _lib.load()
    d = prm.dims
    lead = tuple(jac.shape[:-2])
    rows = 1
    for v in lead:
        rows *= int(v)

    def flat(t, w):
        if t is None:
            return None
        t = _f32(t, 'residual_eval input').reshape(rows, w).contiguous()
        return t

    jac_f = flat(jac, d * d)
    u_f, lap_f, pg_f = flat(u, d), flat(lap, d * d), flat(p_grad, d)
    zone_f, dc_f, fc_f = flat(zone, 1), flat(dcoef, d), flat(fcoef, d)
    mom = torch.empty((rows, d), dtype=torch.float32, device=jac.device) if want_momentum else None
    div = torch.empty((rows,), dtype=torch.float32, device=jac.device) if want_div else None
    _lib.launches += 1
    check(lib.pcfd_residual_eval(_ptr(u_f), jac_f.data_ptr(), _ptr(lap_f), _ptr(pg_f), _ptr(zone_f), _ptr(dc_f), _ptr(fc_f),
                                 rows, C.byref(prm), _ptr(mom), _ptr(div), _stream()), 'pcfd_residual_eval')
    return (mom.reshape(*lead, d) if mom is not None else None, div.reshape(*lead) if div is not None else None)


def mean_squares(x: Tensor) -> Tensor:
    """(..., C) -> (C,) mean of squares over all leading dimensions; a 1-D / 0-D trailing shape counts as one column."""
    lib = _lib.load()
    x = _f32(x, 'mean_squares input')
    cols = int(x.shape[-1]) if x.dim() >= 2 else 1
    xf = x.reshape(-1, cols).contiguous()
    out = torch.empty(cols, dtype=torch.float32, device=x.device)
    _lib.launches += 1
    check(lib.pcfd_mean_squares(xf.data_ptr(), xf.shape[0], cols, out.data_ptr(), _stream()), 'pcfd_mean_squares')
    return out


def relobralo_update(losses: Tensor, n: int, init_losses: Tensor, prev_losses: Tensor, lambda_ema: Tensor, step: Tensor,
                     batch_size: int, alpha: float, beta: float, tau: float, eps: float, seed: int,
                     weights_out: Tensor) -> None:
    lib = _lib.load()
    _lib.launches += 1
    check(lib.pcfd_relobralo_update(losses.data_ptr(), n, init_losses.data_ptr(), prev_losses.data_ptr(),
                                    lambda_ema.data_ptr(), step.data_ptr(), batch_size, alpha, beta, tau, eps, seed,
                                    weights_out.data_ptr(), _stream()), 'pcfd_relobralo_update')


def adam_step(param: Tensor, grad: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, step: Tensor, lr: Tensor,
              beta1: float, beta2: float, eps: float, grad_scale: float = 1.0) -> None:
    lib = _lib.load()
    _lib.launches += 2
    check(lib.pcfd_adam_step(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), step.data_ptr(),
                             lr.data_ptr(), beta1, beta2, eps, grad_scale, param.numel(), _stream()), 'pcfd_adam_step')


def dp_adam_step(peers, exp_avg: Tensor, exp_avg_sq: Tensor, step: Tensor, lr: Tensor, beta1: float, beta2: float,
                 eps: float, grad_scale: float, n: int, epoch: Tensor) -> None:
    """Gradient reduce-scatter + Adam on this rank's slice + parameter all-gather in one kernel (pcfd_dp_adam_step);
    `peers` is a _lib.DpPeers filled from the symmetric-memory handles (common/training.py)."""
    lib = _lib.load()
    _lib.launches += 2
    check(lib.pcfd_dp_adam_step(C.byref(peers), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), step.data_ptr(), lr.data_ptr(),
                                beta1, beta2, eps, grad_scale, n, epoch.data_ptr(), _stream()), 'pcfd_dp_adam_step')


def zero_(t: Tensor) -> None:
    lib = _lib.load()
    _lib.launches += 1
    check(lib.pcfd_zero(t.data_ptr(), t.numel(), _stream()), 'pcfd_zero')


def advance_seed(seed_dev: Tensor) -> None:
    lib = _lib.load()
    _lib.launches += 1
    check(lib.pcfd_advance_seed(seed_dev.data_ptr(), _stream()), 'pcfd_advance_seed')


# ------------------------------------------------------------------------------------------------
# batch ingestion on the device (csrc/ingest.cu; reference dataset/foam_dataset.py:83-90, 360-395)
# ------------------------------------------------------------------------------------------------

def sdf_feature(data: Tensor, n_internal: int, pos_col: int, dims: int, region_col: int, sdf_col: int,
                coord_scale: Optional[Tensor] = None) -> None:
    """In place: data (G, N, F)[..., sdf_col] = signed, max-normalised distance to the nearest boundary point."""
    lib = _lib.load()
    data = _f32(data, 'data')
    if data.dim() != 3 or not data.is_contiguous():
        raise _lib.PcfdError('sdf_feature: data must be a contiguous (G, N, F) tensor')
    g, n, f = data.shape
    if coord_scale is not None:
        coord_scale = _f32(coord_scale, 'coord_scale').contiguous()
        if coord_scale.numel() != dims:
            raise _lib.PcfdError('sdf_feature: coord_scale needs one entry per dimension')
    scratch = torch.empty(int(lib.pcfd_sdf_scratch_bytes(g, n)) // 4, dtype=torch.float32, device=data.device)
    _lib.launches += 2
    with _timed('sdf', 4.0 * g * n * (dims + 2)):
      check(lib.pcfd_sdf_feature(data.data_ptr(), g, n, f, n_internal, pos_col, dims, region_col, sdf_col,
                                 _ptr(coord_scale), scratch.data_ptr(), _stream()), 'pcfd_sdf_feature')


def boundary_one_hot(data: Tensor, n_internal: int, boundary_class: Tensor, n_classes: int, col0: int) -> None:
    """In place: data (G, N, F)[..., col0:col0+n_classes] = one-hot patch id of the boundary rows, zero inside."""
    lib = _lib.load()
    data = _f32(data, 'data')
    g, n, f = data.shape
    if not data.is_contiguous():
        raise _lib.PcfdError('boundary_one_hot: data must be contiguous')
    if boundary_class.dtype != torch.int32 or not boundary_class.is_cuda or tuple(boundary_class.shape) != (g, n - n_internal):
        raise _lib.PcfdError('boundary_one_hot: boundary_class must be a CUDA int32 (G, N - n_internal) tensor')
    _lib.launches += 1
    check(lib.pcfd_boundary_one_hot(data.data_ptr(), g, n, f, n_internal, boundary_class.contiguous().data_ptr(),
                                    n_classes, col0, _stream()), 'pcfd_boundary_one_hot')


def gather_blocks(src: Tensor, ids: Tensor) -> Tensor:
    """out[i] = src[ids[i]] along dimension 0 (collate of resident geometries); any 4- or 8-byte dtype."""
    lib = _lib.load()
    if not (src.is_cuda and ids.is_cuda) or ids.dtype != torch.int64 or not src.is_contiguous():
        raise _lib.PcfdError('gather_blocks: src must be a contiguous CUDA tensor and ids a CUDA int64 tensor')
    block_bytes = src[0].numel() * src.element_size()
    out = torch.empty((ids.numel(),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    _lib.launches += 1
    with _timed('collate', 2.0 * ids.numel() * block_bytes):
      check(lib.pcfd_gather_blocks(src.data_ptr(), block_bytes, ids.contiguous().data_ptr(), ids.numel(), out.data_ptr(),
                                   _stream()), 'pcfd_gather_blocks')
    return out


def gather_blocks_multi(srcs, ids: Tensor, outs: Optional[list] = None) -> list:
    """[src[ids] for src in srcs] along dimension 0 in one launch (at most 16 tensors; empty blocks allowed).
    `outs`: preallocated destinations (static buffers of a captured graph)."""
    lib = _lib.load()
    if ids.dtype != torch.int64 or not ids.is_cuda:
        raise _lib.PcfdError('gather_blocks_multi: ids must be a CUDA int64 tensor')
    ids = ids.contiguous()
    given = outs is not None
    outs = list(outs) if given else []
    nbytes = []
    for i, src in enumerate(srcs):
        if not src.is_cuda or not src.is_contiguous():
            raise _lib.PcfdError('gather_blocks_multi: sources must be contiguous CUDA tensors')
        shape = (ids.numel(),) + tuple(src.shape[1:])
        if given:
            o = outs[i]
            if tuple(o.shape) != shape or o.dtype != src.dtype or not o.is_contiguous() or not o.is_cuda:
                raise _lib.PcfdError('gather_blocks_multi: destination does not match its source')
        else:
            outs.append(torch.empty(shape, dtype=src.dtype, device=src.device))
        nbytes.append(src[0].numel() * src.element_size() if src.shape[0] else 0)
    n = len(srcs)
    src_arr = (C.c_void_p * n)(*[s.data_ptr() for s in srcs])
    dst_arr = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
    len_arr = (C.c_int64 * n)(*nbytes)
    _lib.launches += 1
    with _timed('collate', 2.0 * ids.numel() * sum(nbytes)):
      check(lib.pcfd_gather_blocks_multi(src_arr, dst_arr, len_arr, n, ids.data_ptr(), ids.numel(), _stream()),
            'pcfd_gather_blocks_multi')
    return outs


_IDENTITY_IDS: dict = {}


def copy_blocks_multi(srcs, dsts) -> None:
    """dst[i] <- src[i] for up to 16 equally batched tensors in ONE launch (a batch into the static input buffers of
    a captured step: the data tensor and every sub-domain's ids) instead of one copy kernel per tensor."""
    b = srcs[0].shape[0]
    key = (b, srcs[0].device)
    ids = _IDENTITY_IDS.get(key)
    if ids is None:
        ids = _IDENTITY_IDS[key] = torch.arange(b, dtype=torch.int64, device=srcs[0].device)
    gather_blocks_multi(srcs, ids, outs=dsts)


def set_gemm_engine(engine: int) -> None:
    """Host-side switch between the product path (2) and the generic fp32 CUDA-core engine (0) for every jet layer
    issued from this process (tests / scripts/bench_layers.py).  The library has no such state."""
    global FORCE_FFMA
    if engine not in (0, 2):
        raise _lib.PcfdError('jet GEMM engine must be 0 (fp32 FFMA reference engine) or 2 (tcgen05 product path)')
    FORCE_FFMA = engine == 0
