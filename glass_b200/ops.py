"""torch custom ops over the C ABI of libglass_b200.so, plus their autograd formulas.

Layering
  * primitive ops ``torch.ops.glass_b200.*`` -- registered with torch.library for the CUDA dispatch
    key only (a CPU tensor therefore fails loudly: there is no CPU path).  Each one is a thin ctypes
    call into one ``extern "C"`` entry point of include/glass_b200.h on the current CUDA stream;
    outputs and workspaces are torch tensors allocated by the caller (torch caching allocator), so
    everything except csr_build is CUDA-graph capturable.
  * autograd Functions (SpMM, PairLinearMix, GraphNorm, GraphNormCat, Embedding, SegmentPool) that
    compose the primitives into the differentiable building blocks used by glass_b200.models.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import ACT_NONE, AGGR, GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05, POOL, check

__all__ = ["CSRAdj", "build_csr", "to_undirected", "spmm", "spmm_graph_norm", "graph_norm_pool", "graph_norm_pool_cat", "glass_conv", "conv_fusable", "pair_linear_mix", "graph_norm", "graph_norm_cat", "embedding",
           "segment_pool", "segment_pool_batch", "maxzoz", "label_mask", "pad2batch", "inject_keep_masks",
           "set_gemm_path", "launch_count", "reset_launch_count", "manual_seed"]

# ---------------------------------------------------------------------------------------------
# small helpers
# ---------------------------------------------------------------------------------------------
_launches = 0          # kernels launched through this module (bench.py reports it as gpu_launches)
_gemm_path = {"auto": GEMM_AUTO, "simt": GEMM_SIMT, "tcgen05": GEMM_TCGEN05}[
    os.environ.get("GLASS_B200_GEMM", "auto")]


# GLASSConv as one fused chain (operands normalised inside the GEMM loaders) or as separate ops with the
# single-launch GraphNorm; measured numbers in DESIGN.md section 4.  GLASS_B200_CONV_FUSED=0/1 overrides.
_conv_fused = os.environ.get("GLASS_B200_CONV_FUSED", "0") != "0"


def set_conv_fused(on: bool) -> None:
    global _conv_fused
    _conv_fused = bool(on)


def set_gemm_path(name: str) -> None:
    """'auto' | 'simt' | 'tcgen05' -- which kernel family evaluates the label-mixed Linear pairs."""
    global _gemm_path
    _gemm_path = {"auto": GEMM_AUTO, "simt": GEMM_SIMT, "tcgen05": GEMM_TCGEN05}[name]


def launch_count() -> int:
    return _launches


def reset_launch_count() -> None:
    global _launches
    _launches = 0


def _count(n: int) -> None:
    global _launches
    _launches += n


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t: torch.Tensor, dtype, name: str, dim: Optional[int] = None) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"glass_b200: `{name}` is on {t.device}; the hot path has no CPU implementation")
    if t.dtype != dtype:
        raise RuntimeError(f"glass_b200: `{name}` must be {dtype}, got {t.dtype}")
    if dim is not None and t.dim() != dim:
        raise RuntimeError(f"glass_b200: `{name}` must be {dim}-d, got shape {tuple(t.shape)}")
    return t if t.is_contiguous() else t.contiguous()


def _rowmajor(t: torch.Tensor):
    """(tensor usable as a row-major matrix, leading dimension).  Column slices of a contiguous
    matrix are accepted as-is (that is how the JK concat buffer is written without a copy)."""
    if t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]:
        return t, t.stride(0)
    t = t.contiguous()
    return t, t.shape[1]


# ---------------------------------------------------------------------------------------------
# primitive ops (CUDA dispatch key only)
# ---------------------------------------------------------------------------------------------
_L = torch.library.Library("glass_b200", "DEF")


def _define(schema: str, fn):
    _L.define(schema)
    _L.impl(schema.split("(")[0], fn, "CUDA")


def _spmm_csr_(rowptr, col, val, x, y, partial):
    """partial: optional fp64 [2*h, ld] table that receives the per-CTA column sums of y and y^2 (GraphNorm
    statistics from the SpMM epilogue); returns the number of blocks written (0 without `partial`)."""
    lib = _lib.load()
    n_rows, h = y.shape
    nblk = C.c_int(0)
    check(lib.glass_spmm_csr(_p(rowptr), _p(col), _p(val), _p(x), x.stride(0), _p(y), y.stride(0), n_rows,
                             x.shape[0], h, _p(partial), 0 if partial is None else partial.stride(0),
                             C.byref(nblk), _stream()), "spmm_csr")
    _count(1)
    return nblk.value


_define("spmm_csr_(Tensor rowptr, Tensor col, Tensor val, Tensor x, Tensor(a!) y, Tensor(b!)? partial) -> int",
        _spmm_csr_)


def _spmm_csr_planned_(col, val, x, y, item_begin, item_end, item_dst, long_row, long_slot, long_cnt, scratch,
                       partial):
    lib = _lib.load()
    n_rows, h = y.shape
    nblk = C.c_int(0)
    check(lib.glass_spmm_csr_planned(_p(col), _p(val), _p(x), x.stride(0), _p(y), y.stride(0), n_rows, x.shape[0], h,
                                     _p(item_begin), _p(item_end), _p(item_dst), item_begin.numel(), _p(long_row),
                                     _p(long_slot), _p(long_cnt), long_row.numel(), _p(scratch), _p(partial),
                                     0 if partial is None else partial.stride(0), C.byref(nblk), _stream()),
          "spmm_csr_planned")
    _count(2)
    return nblk.value


_define("spmm_csr_planned_(Tensor col, Tensor val, Tensor x, Tensor(a!) y, Tensor item_begin, Tensor item_end, "
        "Tensor item_dst, Tensor long_row, Tensor long_slot, Tensor long_cnt, Tensor(b!) scratch, "
        "Tensor(c!)? partial) -> int", _spmm_csr_planned_)


def _spmm_csr_acc_(rowptr, col, val, x, y, item_begin, item_end, item_dst, long_row, long_slot, long_cnt, scratch,
                   accumulate):
    """y (+)= A x, plain CSR (item_begin None) or row-split plan; see glass_spmm_csr_acc."""
    lib = _lib.load()
    n_rows, h = y.shape
    planned = item_begin is not None
    check(lib.glass_spmm_csr_acc(_p(rowptr), _p(col), _p(val), _p(x), x.stride(0), _p(y), y.stride(0), n_rows, x.shape[0],
                                 h, _p(item_begin), _p(item_end), _p(item_dst), item_begin.numel() if planned else 0,
                                 _p(long_row), _p(long_slot), _p(long_cnt), long_row.numel() if planned else 0,
                                 _p(scratch), 1 if accumulate else 0, _stream()), "spmm_csr_acc")
    _count(2 if planned else 1)


_define("spmm_csr_acc_(Tensor? rowptr, Tensor col, Tensor val, Tensor x, Tensor(a!) y, Tensor? item_begin, "
        "Tensor? item_end, Tensor? item_dst, Tensor? long_row, Tensor? long_slot, Tensor? long_cnt, "
        "Tensor(b!)? scratch, bool accumulate) -> ()", _spmm_csr_acc_)


def _spmm_delta_(rowptr, col, val, mask, delta, base, y, item_begin, item_end, item_dst, long_row, long_slot, long_cnt,
                 scratch, partial):
    """y = base + adj[:, mask] @ delta[mask]  (sparse label correction)."""
    lib = _lib.load()
    n_rows, h = y.shape
    nblk = C.c_int(0)
    planned = item_begin is not None
    check(lib.glass_spmm_delta(_p(rowptr), _p(col), _p(val), _p(mask), _p(delta), delta.stride(0), _p(base),
                               base.stride(0), _p(y), y.stride(0), n_rows, h, _p(item_begin), _p(item_end),
                               _p(item_dst), item_begin.numel() if planned else 0, _p(long_row), _p(long_slot),
                               _p(long_cnt), long_row.numel() if planned else 0, _p(scratch), _p(partial),
                               0 if partial is None else partial.stride(0), C.byref(nblk), _stream()), "spmm_delta")
    _count(2 if planned else 1)
    return nblk.value


_define("spmm_delta_(Tensor? rowptr, Tensor col, Tensor val, Tensor mask, Tensor delta, Tensor base, Tensor(a!) y, "
        "Tensor? item_begin, Tensor? item_end, Tensor? item_dst, Tensor? long_row, Tensor? long_slot, Tensor? long_cnt, "
        "Tensor(b!)? scratch, Tensor(c!)? partial) -> int", _spmm_delta_)


def _pair_fwd_(a1, a2, w0, b0, w1, b1, mask, z_ratio, act, path, out, acts):
    lib = _lib.load()
    n, k1 = a1.shape
    k2 = 0 if a2 is None else a2.shape[1]
    h = w0.shape[0]
    check(lib.glass_pair_linear_mix_fwd(_p(a1), a1.stride(0), k1, _p(a2), 0 if a2 is None else a2.stride(0), k2,
                                        _p(w0), _p(b0), _p(w1), _p(b1), _p(mask), z_ratio, act, _p(out),
                                        out.stride(0), _p(acts), n, h, path, _stream()), "pair_linear_mix_fwd")
    _count(1)


_define("pair_linear_mix_fwd_(Tensor a1, Tensor? a2, Tensor w0, Tensor b0, Tensor w1, Tensor b1, Tensor mask, "
        "float z_ratio, int act, int path, Tensor(a!) out, Tensor(b!)? acts) -> ()", _pair_fwd_)


def _pair_bwd_(dout, acts, a1, a2, w0, w1, mask, z_ratio, act, path, da1, da2, dw0, db0, dw1, db1, workspace):
    lib = _lib.load()
    n, k1 = a1.shape
    k2 = 0 if a2 is None else a2.shape[1]
    h = w0.shape[0]
    check(lib.glass_pair_linear_mix_bwd(
        _p(dout), dout.stride(0), _p(acts), _p(a1), a1.stride(0), k1, _p(a2), 0 if a2 is None else a2.stride(0), k2,
        _p(w0), _p(w1), _p(mask), z_ratio, act, _p(da1), 0 if da1 is None else da1.stride(0), _p(da2),
        0 if da2 is None else da2.stride(0), _p(dw0), _p(db0), _p(dw1), _p(db1), n, h, _p(workspace),
        workspace.numel(), path, _stream()), "pair_linear_mix_bwd")
    _count(3 if (da1 is not None or da2 is not None) else 2)


_define("pair_linear_mix_bwd_(Tensor dout, Tensor? acts, Tensor a1, Tensor? a2, Tensor w0, Tensor w1, Tensor mask, "
        "float z_ratio, int act, int path, Tensor(a!)? da1, Tensor(b!)? da2, Tensor(c!) dw0, Tensor(d!) db0, "
        "Tensor(e!) dw1, Tensor(f!) db1, Tensor(g!) workspace) -> ()", _pair_bwd_)


def _norm_operand(stats, bits, p, act):
    """ctypes view of a row operand that the GEMM loader normalises on the fly (None: plain operand)."""
    if stats is None:
        return None
    return C.byref(_lib.NormOperand(stats.data_ptr(), 0 if bits is None else bits.data_ptr(), float(p), int(act)))


def _pair_fwd_ex_(a1, a2, w0, b0, w1, b1, mask, z_ratio, act, path, out, acts, stats1, bits1, p1, act1, stats2, bits2,
                  p2, act2):
    lib = _lib.load()
    n, k1 = a1.shape
    k2 = 0 if a2 is None else a2.shape[1]
    h = w0.shape[0]
    check(lib.glass_pair_linear_mix_fwd_ex(_p(a1), a1.stride(0), k1, _p(a2), 0 if a2 is None else a2.stride(0), k2,
                                           _p(w0), _p(b0), _p(w1), _p(b1), _p(mask), z_ratio, act, _p(out),
                                           out.stride(0), _p(acts), n, h, path, _norm_operand(stats1, bits1, p1, act1),
                                           _norm_operand(stats2, bits2, p2, act2), _stream()), "pair_linear_mix_fwd_ex")
    _count(1)


_define("pair_linear_mix_fwd_ex_(Tensor a1, Tensor? a2, Tensor w0, Tensor b0, Tensor w1, Tensor b1, Tensor mask, "
        "float z_ratio, int act, int path, Tensor(a!) out, Tensor(b!)? acts, Tensor? stats1, Tensor? bits1, float p1, "
        "int act1, Tensor? stats2, Tensor? bits2, float p2, int act2) -> ()", _pair_fwd_ex_)


def _pair_bwd_ex_(dout, acts, a1, a2, w0, w1, mask, z_ratio, act, path, da1, da2, dw0, db0, dw1, db1, workspace,
                  stats1, bits1, p1, act1, stats2, bits2, p2, act2, acc1, acc2):
    lib = _lib.load()
    n, k1 = a1.shape
    k2 = 0 if a2 is None else a2.shape[1]
    h = w0.shape[0]
    check(lib.glass_pair_linear_mix_bwd_ex(
        _p(dout), dout.stride(0), _p(acts), _p(a1), a1.stride(0), k1, _p(a2), 0 if a2 is None else a2.stride(0), k2,
        _p(w0), _p(w1), _p(mask), z_ratio, act, _p(da1), 0 if da1 is None else da1.stride(0), _p(da2),
        0 if da2 is None else da2.stride(0), _p(dw0), _p(db0), _p(dw1), _p(db1), n, h, _p(workspace),
        workspace.numel(), path, _norm_operand(stats1, bits1, p1, act1), _norm_operand(stats2, bits2, p2, act2),
        int(acc1), int(acc2), _stream()), "pair_linear_mix_bwd_ex")
    _count(3 if (da1 is not None or da2 is not None) else 2)


_define("pair_linear_mix_bwd_ex_(Tensor dout, Tensor? acts, Tensor a1, Tensor? a2, Tensor w0, Tensor w1, Tensor mask, "
        "float z_ratio, int act, int path, Tensor(a!)? da1, Tensor(b!)? da2, Tensor(c!) dw0, Tensor(d!) db0, "
        "Tensor(e!) dw1, Tensor(f!) db1, Tensor(g!) workspace, Tensor? stats1, Tensor? bits1, float p1, int act1, "
        "Tensor? stats2, Tensor? bits2, float p2, int act2, int acc1, int acc2) -> ()", _pair_bwd_ex_)


def _gn_fwd_(x, weight, bias, mean_scale, eps, act, keep, drop_p, rng, bits, out, stats, workspace):
    lib = _lib.load()
    n, c = x.shape
    check(lib.glass_graphnorm_fwd(_p(x), x.stride(0), _p(weight), _p(bias), _p(mean_scale), eps, act, _p(keep),
                                  drop_p, _p(rng), _p(bits), _p(out), out.stride(0), _p(stats), n, c, _p(workspace),
                                  workspace.numel(), _stream()), "graphnorm_fwd")
    _count(lib.glass_graphnorm_launches(n, c))


_define("graphnorm_fwd_(Tensor x, Tensor weight, Tensor bias, Tensor mean_scale, float eps, int act, Tensor? keep, "
        "float drop_p, Tensor(d!)? rng, Tensor(e!)? bits, Tensor(a!) out, Tensor(b!) stats, Tensor(c!) workspace) -> ()",
        _gn_fwd_)


def _gn_bwd_(dout, x, weight, mean_scale, stats, act, keep, drop_p, rng, bits, dx, dweight, dbias, dmean_scale,
             workspace):
    lib = _lib.load()
    n, c = x.shape
    check(lib.glass_graphnorm_bwd(_p(dout), dout.stride(0), _p(x), x.stride(0), _p(weight), _p(mean_scale),
                                  _p(stats), act, _p(keep), drop_p, _p(rng), _p(bits), _p(dx), dx.stride(0),
                                  _p(dweight), _p(dbias), _p(dmean_scale), n, c, _p(workspace), workspace.numel(),
                                  _stream()), "graphnorm_bwd")
    _count(lib.glass_graphnorm_launches(n, c))


_define("graphnorm_bwd_(Tensor dout, Tensor x, Tensor weight, Tensor mean_scale, Tensor stats, int act, Tensor? keep, "
        "float drop_p, Tensor? rng, Tensor? bits, Tensor(a!) dx, Tensor(b!) dweight, Tensor(c!) dbias, "
        "Tensor(d!) dmean_scale, Tensor(e!) workspace) -> ()", _gn_bwd_)


def _gn_stats_(partial, nblk, n, weight, bias, mean_scale, eps, keep, drop_p, rng, bits, stats):
    lib = _lib.load()
    c = weight.numel()
    check(lib.glass_graphnorm_stats(_p(partial), nblk, partial.stride(0), _p(weight), _p(bias), _p(mean_scale), eps,
                                    _p(keep), drop_p, _p(rng), _p(bits), _p(stats), n, c, _stream()), "graphnorm_stats")
    _count(1)


_define("graphnorm_stats_(Tensor partial, int nblk, int n, Tensor weight, Tensor bias, Tensor mean_scale, float eps, "
        "Tensor? keep, float drop_p, Tensor(c!)? rng, Tensor(b!)? bits, Tensor(a!) stats) -> ()", _gn_stats_)


def _gn_apply_(x, stats, act, keep, drop_p, bits, out):
    lib = _lib.load()
    n, c = x.shape
    check(lib.glass_graphnorm_apply(_p(x), x.stride(0), _p(stats), act, _p(keep), drop_p, _p(bits), _p(out),
                                    out.stride(0), n, c, _stream()), "graphnorm_apply")
    _count(1)


_define("graphnorm_apply_(Tensor x, Tensor stats, int act, Tensor? keep, float drop_p, Tensor? bits, Tensor(a!) out) -> ()",
        _gn_apply_)


def _gn_bwd_from_sums_(partial, nblk, u, x, weight, mean_scale, stats, dx, dweight, dbias, dmean_scale, workspace):
    lib = _lib.load()
    n, c = x.shape
    check(lib.glass_graphnorm_bwd_from_sums(_p(partial), nblk, partial.stride(0), _p(u), u.stride(0), _p(x),
                                            x.stride(0), _p(weight), _p(mean_scale), _p(stats), _p(dx), dx.stride(0),
                                            _p(dweight), _p(dbias), _p(dmean_scale), n, c, _p(workspace),
                                            workspace.numel(), _stream()), "graphnorm_bwd_from_sums")
    _count(2)


_define("graphnorm_bwd_from_sums_(Tensor partial, int nblk, Tensor u, Tensor x, Tensor weight, Tensor mean_scale, "
        "Tensor stats, Tensor(a!) dx, Tensor(b!) dweight, Tensor(c!) dbias, Tensor(d!) dmean_scale, "
        "Tensor(e!) workspace) -> ()", _gn_bwd_from_sums_)


def _emb_fwd_(table, ids, out):
    lib = _lib.load()
    check(lib.glass_embedding_fwd(_p(table), _p(ids), _p(out), out.stride(0), ids.numel(), table.shape[0],
                                  table.shape[1], _stream()), "embedding_fwd")
    _count(1)


def _emb_bwd_(dout, ids, dtable):
    lib = _lib.load()
    check(lib.glass_embedding_bwd(_p(dout), dout.stride(0), _p(ids), _p(dtable), ids.numel(), dtable.shape[0],
                                  dtable.shape[1], _stream()), "embedding_bwd")
    _count(1)


_define("embedding_fwd_(Tensor table, Tensor ids, Tensor(a!) out) -> ()", _emb_fwd_)
_define("embedding_bwd_(Tensor dout, Tensor ids, Tensor(a!) dtable) -> ()", _emb_bwd_)


def _pool_fwd_(emb, pos, mode, out, cnt, argmax):
    lib = _lib.load()
    b, lmax = pos.shape
    check(lib.glass_segment_pool_fwd(_p(emb), emb.stride(0), _p(pos), b, lmax, mode, _p(out), out.stride(0), _p(cnt),
                                     _p(argmax), emb.shape[1], emb.shape[0], _stream()), "segment_pool_fwd")
    _count(1)


def _norm_pool_fwd_(x, stats, pos, mode, out, cnt, ysum):
    lib = _lib.load()
    b, lmax = pos.shape
    check(lib.glass_norm_pool_fwd(_p(x), x.stride(0), _p(stats), _p(pos), b, lmax, mode, _p(out), out.stride(0), _p(cnt),
                                  _p(ysum), x.shape[1], x.shape[0], _stream()), "norm_pool_fwd")
    _count(1)


def _norm_pool_bwd_(dout, pos, mode, cnt, ysum, x, stats, weight, mean_scale, dx, dw, db, dms, scratch):
    lib = _lib.load()
    b, lmax = pos.shape
    check(lib.glass_norm_pool_bwd(_p(dout), dout.stride(0), _p(pos), b, lmax, mode, _p(cnt), _p(ysum), _p(x), x.stride(0),
                                  _p(stats), _p(weight), _p(mean_scale), _p(dx), dx.stride(0), _p(dw), _p(db), _p(dms),
                                  x.shape[1], x.shape[0], _p(scratch), scratch.numel(), _stream()), "norm_pool_bwd")
    _count(2)


def _pool_bwd_(dout, pos, mode, cnt, argmax, demb, mark):
    """mark (uint8 [n_node] scratch): deterministic node-centric variant that writes every row of demb."""
    lib = _lib.load()
    b, lmax = pos.shape
    check(lib.glass_segment_pool_bwd(_p(dout), dout.stride(0), _p(pos), b, lmax, mode, _p(cnt), _p(argmax), _p(demb),
                                     demb.stride(0), demb.shape[1], demb.shape[0], _p(mark),
                                     0 if mark is None else mark.numel(), _stream()), "segment_pool_bwd")
    _count(1 if mark is None else 2)


_define("segment_pool_fwd_(Tensor emb, Tensor pos, int mode, Tensor(a!) out, Tensor(b!) cnt, Tensor(c!)? argmax) -> ()",
        _pool_fwd_)
_define("segment_pool_bwd_(Tensor dout, Tensor pos, int mode, Tensor cnt, Tensor? argmax, Tensor(a!) demb, "
        "Tensor(b!)? mark) -> ()", _pool_bwd_)


_define("norm_pool_fwd_(Tensor x, Tensor stats, Tensor pos, int mode, Tensor(a!) out, Tensor(b!) cnt, Tensor(c!) ysum) -> ()",
        _norm_pool_fwd_)
_define("norm_pool_bwd_(Tensor dout, Tensor pos, int mode, Tensor cnt, Tensor ysum, Tensor x, Tensor stats, Tensor weight, "
        "Tensor mean_scale, Tensor(a!) dx, Tensor(b!) dw, Tensor(c!) db, Tensor(d!) dms, Tensor(e!) scratch) -> ()",
        _norm_pool_bwd_)


def _pool_batch_fwd_(x, batch, n_seg, mode, out, cnt, argmax):
    lib = _lib.load()
    check(lib.glass_segment_pool_batch_fwd(_p(x), x.stride(0), _p(batch), x.shape[0], n_seg, mode, _p(out),
                                           out.stride(0), _p(cnt), _p(argmax), x.shape[1], _stream()),
          "segment_pool_batch_fwd")
    _count(1)


def _pool_batch_bwd_(dout, batch, n_seg, mode, cnt, argmax, dx):
    lib = _lib.load()
    check(lib.glass_segment_pool_batch_bwd(_p(dout), dout.stride(0), _p(batch), dx.shape[0], n_seg, mode, _p(cnt),
                                           _p(argmax), _p(dx), dx.stride(0), dx.shape[1], _stream()),
          "segment_pool_batch_bwd")
    _count(1)


_define("segment_pool_batch_fwd_(Tensor x, Tensor batch, int n_seg, int mode, Tensor(a!) out, Tensor(b!) cnt, "
        "Tensor(c!)? argmax) -> ()", _pool_batch_fwd_)
_define("segment_pool_batch_bwd_(Tensor dout, Tensor batch, int n_seg, int mode, Tensor cnt, Tensor? argmax, "
        "Tensor(a!) dx) -> ()", _pool_batch_bwd_)


def _maxzoz_(pos, z, mask):
    lib = _lib.load()
    check(lib.glass_maxzoz(_p(pos), pos.numel(), _p(z), _p(mask), z.numel(), _stream()), "maxzoz")
    _count(1)


def _label_mask_(z, mask):
    lib = _lib.load()
    check(lib.glass_label_mask(_p(z), _p(mask), z.numel(), _stream()), "label_mask")
    _count(1)


def _pad2batch_(pad, batch_out, pos_out, n_valid):
    lib = _lib.load()
    check(lib.glass_pad2batch(_p(pad), pad.shape[0], pad.shape[1], _p(batch_out), _p(pos_out), _p(n_valid),
                              _stream()), "pad2batch")
    _count(1)


_define("maxzoz_(Tensor pos, Tensor(a!) z, Tensor(b!)? mask) -> ()", _maxzoz_)
_define("label_mask_(Tensor z, Tensor(a!) mask) -> ()", _label_mask_)
_define("pad2batch_(Tensor pad, Tensor(a!) batch_out, Tensor(b!) pos_out, Tensor(c!) n_valid) -> ()", _pad2batch_)

_ops = torch.ops.glass_b200


# ---------------------------------------------------------------------------------------------
# buildAdj -> CSR
# ---------------------------------------------------------------------------------------------
SPLIT_ROW_LEN = int(os.environ.get("GLASS_B200_SPLIT_ROW_LEN", "0"))     # 0: chosen per graph


def split_row_len(n_rows: int, nnz: int) -> int:
    """Longest row a single lane group reduces.  Graphs that fill the machine are throughput bound: only
    hub rows (> 512 entries) are split.  Below about two waves of lanes (same rule as the SpMM dispatcher)
    the longest row IS the kernel time, so rows are cut at twice the average degree (>= 64)."""
    if SPLIT_ROW_LEN:
        return SPLIT_ROW_LEN
    if n_rows * 32 > 2 * _lib.load().glass_sm_count() * 2048:
        return 512
    target = max(64, 2 * nnz // max(n_rows, 1))
    return min(512, 1 << (target - 1).bit_length())


class RowSplitPlan:
    """Work items for a skewed CSR: rows longer than `max_len` (split_row_len) entries are cut into chunks that are
    reduced by separate lane groups (heavy items first) and summed in chunk order afterwards."""

    def __init__(self, rowptr: torch.Tensor, max_len: int):
        lib = _lib.load()
        n_rows = rowptr.numel() - 1
        sizes = [C.c_int64(0) for _ in range(3)]
        check(lib.glass_spmm_plan_size(_p(rowptr), n_rows, max_len, *[C.byref(v) for v in sizes], _stream()),
              "spmm_plan_size")
        self.n_items, self.n_long, self.n_slots = (v.value for v in sizes)
        i32 = dict(dtype=torch.int32, device=rowptr.device)
        self.item_begin = torch.empty(self.n_items, **i32)
        self.item_end = torch.empty(self.n_items, **i32)
        self.item_dst = torch.empty(self.n_items, **i32)
        self.long_row = torch.empty(self.n_long, **i32)
        self.long_slot = torch.empty(self.n_long, **i32)
        self.long_cnt = torch.empty(self.n_long, **i32)
        if self.n_long:
            check(lib.glass_spmm_plan_build(_p(rowptr), n_rows, max_len, _p(self.item_begin), _p(self.item_end),
                                            _p(self.item_dst), _p(self.long_row), _p(self.long_slot),
                                            _p(self.long_cnt), _stream()), "spmm_plan_build")


class CSRAdj:
    """Normalised adjacency as CSR plus the CSR of its transpose (what buildAdj returns here).

    Stands in for the sparse COO tensor of impl/models.py:83-111: supports ``adj @ x``, ``.shape``
    and, for inspection, ``indices()`` / ``values()`` in coalesced COO form."""

    def __init__(self, n, rowptr, col, val, rowptr_t, col_t, val_t, deg, aggr, plan=None, plan_t=None):
        self.n, self.aggr = n, aggr
        self.rowptr, self.col, self.val = rowptr, col, val
        self.rowptr_t, self.col_t, self.val_t = rowptr_t, col_t, val_t
        self.deg = deg
        self.shape = (n, n)
        self.plan, self.plan_t = plan, plan_t     # RowSplitPlan or None (no row long enough to split)

    def make_plans(self, max_len: int = None):
        max_len = split_row_len(self.n, self.col.numel()) if max_len is None else max_len
        self.plan = self.plan_t = None
        if self.n and self.col.numel():
            p = RowSplitPlan(self.rowptr, max_len)
            self.plan = p if p.n_long else None
            p = RowSplitPlan(self.rowptr_t, max_len)
            self.plan_t = p if p.n_long else None
        return self

    @property
    def nnz(self) -> int:
        return self.col.numel()

    def indices(self) -> torch.Tensor:
        counts = (self.rowptr[1:] - self.rowptr[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(self.n, device=self.col.device), counts)
        return torch.stack((rows, self.col.long()))

    def values(self) -> torch.Tensor:
        return self.val

    def coalesce(self):
        return self

    def t(self):
        return CSRAdj(self.n, self.rowptr_t, self.col_t, self.val_t, self.rowptr, self.col, self.val, self.deg,
                      self.aggr, self.plan_t, self.plan)

    def __matmul__(self, x):
        return spmm(self, x)


def build_csr(edge_index: torch.Tensor, edge_weight: torch.Tensor, n_node: int, aggr: str) -> CSRAdj:
    """GPU restatement of buildAdj (impl/models.py:83-111); raises NotImplementedError for an unknown aggr."""
    if aggr not in AGGR:
        raise NotImplementedError(aggr)  # impl/models.py:110-111
    lib = _lib.load()
    ei = _req(edge_index, torch.int64, "edge_index", 2)
    ew = _req(edge_weight, torch.float32, "edge_weight", 1)
    nnz = ei.shape[1]
    dev = ei.device
    with torch.cuda.device(dev):
        ws_bytes = lib.glass_csr_build_workspace_bytes(nnz, n_node)
        if ws_bytes == 0:
            check(-1, "csr_build_workspace_bytes")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        rowptr, rowptr_t = torch.empty(n_node + 1, **i32), torch.empty(n_node + 1, **i32)
        col, col_t = torch.empty(nnz, **i32), torch.empty(nnz, **i32)
        val, val_t = torch.empty(nnz, **f32), torch.empty(nnz, **f32)
        deg = torch.empty(n_node, **f32)
        nnz_out = C.c_int64(0)
        check(lib.glass_csr_build(_p(ei), _p(ew), nnz, n_node, AGGR[aggr], _p(rowptr), _p(col), _p(val),
                                  _p(rowptr_t), _p(col_t), _p(val_t), _p(deg), C.byref(nnz_out), _p(ws), ws_bytes,
                                  _stream()), "csr_build")
        _count(8)
    m = nnz_out.value
    if m != nnz:  # duplicates were merged
        col, val, col_t, val_t = col[:m].clone(), val[:m].clone(), col_t[:m].clone(), val_t[:m].clone()
    return CSRAdj(n_node, rowptr, col, val, rowptr_t, col_t, val_t, deg, aggr).make_plans()


def to_undirected(edge_index: torch.Tensor, edge_weight: torch.Tensor, n_node: int):
    """Device-side PyG `to_undirected` (reference datasets.py:68-71): (edge_index, edge_weight) symmetrised, sorted by
    (row, col), duplicate weights added; an input that already is undirected and duplicate-free comes back as is."""
    lib = _lib.load()
    ei = _req(edge_index, torch.int64, "edge_index", 2)
    ew = _req(edge_weight, torch.float32, "edge_weight", 1)
    nnz, dev = ei.shape[1], ei.device
    if nnz == 0:
        return ei, ew
    with torch.cuda.device(dev):
        ws_bytes = lib.glass_to_undirected_workspace_bytes(nnz)
        if ws_bytes == 0:
            check(-1, "to_undirected_workspace_bytes")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        out_i = torch.empty((2, 2 * nnz), dtype=torch.int64, device=dev)
        out_w = torch.empty(2 * nnz, dtype=torch.float32, device=dev)
        m, already = C.c_int64(0), C.c_int(0)
        check(lib.glass_to_undirected(_p(ei), _p(ew), nnz, n_node, _p(out_i), _p(out_w), C.byref(m), C.byref(already),
                                      _p(ws), ws_bytes, _stream()), "to_undirected")
        _count(6)
    if already.value:
        return ei, ew
    return out_i[:, :m.value].contiguous(), out_w[:m.value].contiguous()


# ---------------------------------------------------------------------------------------------
# autograd building blocks
# ---------------------------------------------------------------------------------------------
def _run_spmm(rowptr, col, val, plan, x, y, partial=None, accumulate: bool = False) -> int:
    """y = A x (accumulate: y += A x, no statistics).  `plan`: RowSplitPlan of this CSR or None."""
    if accumulate:
        if plan is None:
            _ops.spmm_csr_acc_(rowptr, col, val, x, y, None, None, None, None, None, None, None, True)
        else:
            scratch = torch.empty((plan.n_slots, y.shape[1]), dtype=torch.float32, device=y.device)
            _ops.spmm_csr_acc_(None, col, val, x, y, plan.item_begin, plan.item_end, plan.item_dst, plan.long_row,
                               plan.long_slot, plan.long_cnt, scratch, True)
        return 0
    if plan is None:
        return _ops.spmm_csr_(rowptr, col, val, x, y, partial)
    scratch = torch.empty((plan.n_slots, y.shape[1]), dtype=torch.float32, device=y.device)
    return _ops.spmm_csr_planned_(col, val, x, y, plan.item_begin, plan.item_end, plan.item_dst, plan.long_row,
                                  plan.long_slot, plan.long_cnt, scratch, partial)


def graphnorm_partials(x: torch.Tensor, partial: torch.Tensor) -> int:
    """Per-block fp64 column sums of x and x^2 of THIS rank's rows (row-partitioned GraphNorm, phase 1)."""
    lib = _lib.load()
    nblk = C.c_int(0)
    check(lib.glass_graphnorm_partials(_p(x), x.stride(0), x.shape[0], x.shape[1], _p(partial), partial.stride(0),
                                       C.byref(nblk), _stream()), "graphnorm_partials")
    _count(1)
    return nblk.value


def graphnorm_bwd_partials(dout, x, stats, act, keep, drop_p, bits, partial) -> int:
    lib = _lib.load()
    nblk = C.c_int(0)
    check(lib.glass_graphnorm_bwd_partials(_p(dout), dout.stride(0), _p(x), x.stride(0), _p(stats), act, _p(keep),
                                           drop_p, _p(bits), x.shape[0], x.shape[1], _p(partial), partial.stride(0),
                                           C.byref(nblk), _stream()), "graphnorm_bwd_partials")
    _count(1)
    return nblk.value


def graphnorm_bwd_finish(total, n_total, dout, x, weight, mean_scale, stats, act, keep, drop_p, bits, dx, dw, db, da):
    """Row-partitioned GraphNorm backward, phase 2: `total` [2c, 1] = S1 | S2 summed over all ranks."""
    lib = _lib.load()
    n, c = x.shape
    ws = torch.empty(lib.glass_graphnorm_workspace_bytes(n, c), dtype=torch.uint8, device=x.device)
    check(lib.glass_graphnorm_bwd_finish(_p(total), 1, total.stride(0), n_total, _p(dout), dout.stride(0), _p(x),
                                         x.stride(0), _p(weight), _p(mean_scale), _p(stats), act, _p(keep), drop_p,
                                         _p(bits), _p(dx), dx.stride(0), _p(dw), _p(db), _p(da), n, c, _p(ws),
                                         ws.numel(), _stream()), "graphnorm_bwd_finish")
    _count(2)


def spmm_delta(adj: "CSRAdj", mask: torch.Tensor, delta: torch.Tensor, base: torch.Tensor, y: torch.Tensor,
               partial: Optional[torch.Tensor] = None) -> int:
    """y = base + adj[:, mask] @ delta[mask] (no autograd: evaluation only); returns the statistics block count."""
    plan = adj.plan
    if plan is None:
        return _ops.spmm_delta_(adj.rowptr, adj.col, adj.val, mask, delta, base, y, None, None, None, None, None, None,
                                None, partial)
    scratch = torch.empty((plan.n_slots, y.shape[1]), dtype=torch.float32, device=y.device)
    return _ops.spmm_delta_(None, adj.col, adj.val, mask, delta, base, y, plan.item_begin, plan.item_end, plan.item_dst,
                            plan.long_row, plan.long_slot, plan.long_cnt, scratch, partial)


def _stats_table(c: int, device) -> torch.Tensor:
    """fp64 [2*c, ld] table for per-CTA partial column sums produced by a kernel epilogue."""
    return torch.empty((2 * c, _lib.load().glass_spmm_stats_ld()), dtype=torch.float64, device=device)


class _SpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, adj: CSRAdj):
        x, _ = _rowmajor(_req(x, torch.float32, "x", 2))
        if x.shape[0] != adj.n:
            raise RuntimeError(f"spmm: x has {x.shape[0]} rows but the adjacency is {adj.n} x {adj.n}")
        y = torch.empty((adj.n, x.shape[1]), dtype=torch.float32, device=x.device)
        _run_spmm(adj.rowptr, adj.col, adj.val, adj.plan, x, y)
        ctx.adj = adj
        return y

    @staticmethod
    def backward(ctx, gy):
        adj = ctx.adj
        gy, _ = _rowmajor(gy)
        gx = torch.empty_like(gy, memory_format=torch.contiguous_format)
        _run_spmm(adj.rowptr_t, adj.col_t, adj.val_t, adj.plan_t, gy, gx)  # dX = A^T dY
        return gx, None


def spmm(adj: CSRAdj, x: torch.Tensor) -> torch.Tensor:
    """``adj @ x`` (impl/models.py:164)."""
    return _SpMM.apply(x, adj)


class _PairLinearMix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a1, a2, w0, b0, w1, b1, mask, z_ratio, act, path):
        a1, _ = _rowmajor(_req(a1, torch.float32, "a1", 2))
        if a2 is not None:
            a2, _ = _rowmajor(_req(a2, torch.float32, "a2", 2))
        w0, b0 = _req(w0, torch.float32, "w0", 2), _req(b0, torch.float32, "b0", 1)
        w1, b1 = _req(w1, torch.float32, "w1", 2), _req(b1, torch.float32, "b1", 1)
        mask = _req(mask, torch.uint8, "mask", 1)
        n, h = a1.shape[0], w0.shape[0]
        k = a1.shape[1] + (0 if a2 is None else a2.shape[1])
        if w0.shape != (h, k) or w1.shape != (h, k) or mask.shape[0] != n:
            raise RuntimeError(f"pair_linear_mix: shape mismatch a={n}x{k} w0={tuple(w0.shape)} w1={tuple(w1.shape)}")
        need_grad = any(ctx.needs_input_grad[:6])
        out = torch.empty((n, h), dtype=torch.float32, device=a1.device)
        acts = (torch.empty((n, 2 * h), dtype=torch.float32, device=a1.device)
                if (need_grad and act != ACT_NONE) else None)
        _ops.pair_linear_mix_fwd_(a1, a2, w0, b0, w1, b1, mask, float(z_ratio), act, path, out, acts)
        ctx.save_for_backward(a1, a2, w0, w1, mask, acts)
        ctx.cfg = (float(z_ratio), act, path)
        return out

    @staticmethod
    def backward(ctx, dout):
        a1, a2, w0, w1, mask, acts = ctx.saved_tensors
        z_ratio, act, path = ctx.cfg
        dout, _ = _rowmajor(dout)
        n, h = dout.shape
        k = w0.shape[1]
        dev = dout.device
        da1 = torch.empty_like(a1, memory_format=torch.contiguous_format) if ctx.needs_input_grad[0] else None
        da2 = (torch.empty_like(a2, memory_format=torch.contiguous_format)
               if (a2 is not None and ctx.needs_input_grad[1]) else None)
        dw0, dw1 = torch.empty_like(w0), torch.empty_like(w1)
        db0 = torch.empty(h, dtype=torch.float32, device=dev)
        db1 = torch.empty(h, dtype=torch.float32, device=dev)
        ws = torch.empty(_lib.load().glass_pair_linear_mix_bwd_workspace_bytes(n, h, k), dtype=torch.uint8, device=dev)
        _ops.pair_linear_mix_bwd_(dout, acts, a1, a2, w0, w1, mask, z_ratio, act, path, da1, da2, dw0, db0, dw1, db1,
                                  ws)
        return da1, da2, dw0, db0, dw1, db1, None, None, None, None


def pair_linear_mix(a1, a2, w0, b0, w1, b1, mask, z_ratio: float, act: int, path: Optional[int] = None):
    """where(mask, z*p1+(1-z)*p0, z*p0+(1-z)*p1) with p_i = act([a1|a2] W_i^T + b_i)
    (impl/models.py:158-162 and :167-173)."""
    return _PairLinearMix.apply(a1, a2, w0, b0, w1, b1, mask, z_ratio, act, _gemm_path if path is None else path)


_keep_queue: Optional[List[torch.Tensor]] = None


@contextlib.contextmanager
def inject_keep_masks(masks: Sequence[torch.Tensor]):
    """Testing hook: dropout keep-masks (uint8 [n, c]) consumed in call order instead of fresh draws,
    so a train()-mode pass can be compared element-wise with the oracle given the same masks."""
    global _keep_queue
    _keep_queue = list(masks)
    try:
        yield
    finally:
        _keep_queue = None


def _injected_keep(n, c):
    """Next injected keep mask, or None when the kernels should draw the bits themselves."""
    if _keep_queue is None:
        return None
    m = _keep_queue.pop(0)
    assert m.shape == (n, c), (m.shape, (n, c))
    return _req(m, torch.uint8, "keep", 2)


_rng_states = {}
_rng_seed = None        # set by manual_seed(); devices whose state does not exist yet start from it


def manual_seed(seed: int) -> None:
    """Re-seed the in-kernel dropout generator of every device (counter back to 0)."""
    global _rng_seed
    _rng_seed = seed & ((1 << 63) - 1)
    for t in _rng_states.values():
        t.copy_(torch.tensor([seed & ((1 << 63) - 1), 0, 0, 0], dtype=torch.int64))


def _rng_state(device) -> torch.Tensor:
    """{seed, call counter, ticket, pad} of the dropout generator on `device`.  The seed is torch's current initial seed
    (read, not drawn: torch's own random stream -- and with it the reference-identical DataLoader shuffles --
    is left untouched), so torch.manual_seed(...) before the first training step makes runs repeatable."""
    key = (device.type, device.index)
    t = _rng_states.get(key)
    if t is None:
        seed = _rng_seed if _rng_seed is not None else int(torch.initial_seed()) & ((1 << 63) - 1)
        t = _rng_states[key] = torch.tensor([seed, 0, 0, 0], dtype=torch.int64, device=device)
    return t


def _gn_workspace(n, c, device):
    return torch.empty(_lib.load().glass_graphnorm_workspace_bytes(n, c), dtype=torch.uint8, device=device)


def pack_keep_bits(keep: torch.Tensor) -> torch.Tensor:
    """uint8 keep mask [n, c] -> packed int32 words, bit (r*c + col) % 32 of word (r*c + col) // 32 (the layout the
    finalize kernel draws and every fused consumer reads)."""
    flat = keep.reshape(-1).to(torch.int64)
    pad = (-flat.numel()) % 32
    if pad:
        flat = torch.cat((flat, flat.new_zeros(pad)))
    w = (flat.view(-1, 32) << torch.arange(32, device=flat.device, dtype=torch.int64)).sum(dim=1)
    return ((w + (1 << 31)) % (1 << 32) - (1 << 31)).to(torch.int32)


def _dropout_source(n: int, c: int, p: float, training: bool, device):
    """(drop_p, keep, rng, bits) for one dropout site.  Generator mode: `bits` is an empty buffer that the
    finalize kernel fills with this call's keep bits.  Injected masks (tests): the explicit uint8 mask for the
    GraphNorm kernels plus its packed form for the fused consumers."""
    if not training or p <= 0.0:
        return 0.0, None, None, None
    if p >= 1.0:
        raise RuntimeError("dropout p must be < 1")
    keep = _injected_keep(n, c)
    if keep is not None:
        return float(p), keep, None, pack_keep_bits(keep)
    words = _lib.load().glass_dropout_bits_bytes(n, c) // 4
    return float(p), None, _rng_state(device), torch.empty(words, dtype=torch.int32, device=device)


_grad_sinks = {}


def register_grad_buffer(param: torch.Tensor, buffer: torch.Tensor) -> None:
    """Have the backward pass write the gradient of `param` straight into `buffer` (same shape, contiguous) when it
    is produced by a GraphNorm backward -- the N x H embedding table enters the model through emb_gn, and the
    data-parallel exchange (glass_b200/dp.py) wants its gradient in peer-visible memory without a 14.7 MB copy."""
    if buffer.shape != param.shape or not buffer.is_contiguous() or buffer.dtype != param.dtype:
        raise ValueError("gradient buffer must match the parameter (shape, dtype, contiguous)")
    import weakref
    for k in [k for k, (ref, _) in _grad_sinks.items() if ref() is None]:      # parameters that no longer exist
        del _grad_sinks[k]
    _grad_sinks[param.data_ptr()] = (weakref.ref(param), buffer)


def unregister_grad_buffer(buffer: torch.Tensor) -> None:
    for k in [k for k, (_, b) in _grad_sinks.items() if b is buffer]:
        del _grad_sinks[k]


def _grad_sink_for(x: torch.Tensor):
    """The registered gradient buffer of `x`, only while the parameter it was registered for is alive and still owns
    that address (a freed parameter's address is recycled by the caching allocator)."""
    hit = _grad_sinks.get(x.data_ptr())
    if hit is None:
        return None
    param, sink = hit[0](), hit[1]
    if param is None or param.data_ptr() != x.data_ptr() or param.shape != x.shape or sink.device != x.device:
        return None
    return sink


class _GraphNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, mean_scale, eps, act, p, training):
        x, _ = _rowmajor(_req(x, torch.float32, "x", 2))
        weight, bias = _req(weight, torch.float32, "weight", 1), _req(bias, torch.float32, "bias", 1)
        mean_scale = _req(mean_scale, torch.float32, "mean_scale", 1)
        n, c = x.shape
        drop_p, keep, rng, bits = _dropout_source(n, c, p, training, x.device)
        out = torch.empty((n, c), dtype=torch.float32, device=x.device)
        stats = torch.empty((6, c), dtype=torch.float32, device=x.device)
        _ops.graphnorm_fwd_(x, weight, bias, mean_scale, float(eps), act, keep, drop_p, rng, bits, out, stats,
                            _gn_workspace(n, c, x.device))
        ctx.save_for_backward(x, weight, mean_scale, stats, keep, rng, bits)
        ctx.cfg = (act, drop_p)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, mean_scale, stats, keep, rng, bits = ctx.saved_tensors
        act, drop_p = ctx.cfg
        dout, _ = _rowmajor(dout)
        n, c = x.shape
        sink = _grad_sink_for(x)
        if sink is not None:
            dx = sink.view(sink.shape)      # a fresh alias: autograd adopts it as .grad without cloning
        else:
            dx = torch.empty((n, c), dtype=torch.float32, device=x.device)
        dw, db, da = torch.empty_like(weight), torch.empty_like(weight), torch.empty_like(weight)
        _ops.graphnorm_bwd_(dout, x, weight, mean_scale, stats, act, keep, drop_p, rng, bits, dx, dw, db, da,
                            _gn_workspace(n, c, x.device))
        return dx, dw, db, da, None, None, None, None


def graph_norm(x, weight, bias, mean_scale, eps: float = 1e-5, act: int = ACT_NONE, p: float = 0.0,
               training: bool = False):
    """dropout(act(GraphNorm(x))) over the whole graph (PyG GraphNorm, batch=None; impl/models.py:165-166,
    249-251, 257-259).  act / dropout are the ops the reference applies right after the norm."""
    return _GraphNorm.apply(x, weight, bias, mean_scale, eps, act, p, training)


class _GraphNormCat(torch.autograd.Function):
    """gns[-1](torch.cat(xs, -1)) of impl/models.py:263-267: GraphNorm is per column, so every xs[l] is
    normalised with its slice of the parameters straight into its column block of the output."""

    @staticmethod
    def forward(ctx, weight, bias, mean_scale, eps, *xs):
        xs = [_rowmajor(_req(t, torch.float32, "xs", 2))[0] for t in xs]
        n = xs[0].shape[0]
        widths = [t.shape[1] for t in xs]
        d = sum(widths)
        weight, bias = _req(weight, torch.float32, "weight", 1), _req(bias, torch.float32, "bias", 1)
        mean_scale = _req(mean_scale, torch.float32, "mean_scale", 1)
        if weight.numel() != d:
            raise RuntimeError(f"graph_norm_cat: {d} columns but {weight.numel()} parameters")
        out = torch.empty((n, d), dtype=torch.float32, device=xs[0].device)
        stats, off = [], 0
        for t, w in zip(xs, widths):
            st = torch.empty((6, w), dtype=torch.float32, device=t.device)
            _ops.graphnorm_fwd_(t, weight[off:off + w], bias[off:off + w], mean_scale[off:off + w], float(eps),
                                ACT_NONE, None, 0.0, None, None, out[:, off:off + w], st,
                                _gn_workspace(n, w, t.device))
            stats.append(st)
            off += w
        ctx.save_for_backward(weight, mean_scale, *xs, *stats)
        ctx.widths = widths
        return out

    @staticmethod
    def backward(ctx, dout):
        widths = ctx.widths
        L = len(widths)
        weight, mean_scale = ctx.saved_tensors[:2]
        xs, stats = ctx.saved_tensors[2:2 + L], ctx.saved_tensors[2 + L:]
        dout, _ = _rowmajor(dout)
        n = dout.shape[0]
        dw, db, da = torch.empty_like(weight), torch.empty_like(weight), torch.empty_like(weight)
        dxs, off = [], 0
        for t, st, w in zip(xs, stats, widths):
            dx = torch.empty((n, w), dtype=torch.float32, device=t.device)
            _ops.graphnorm_bwd_(dout[:, off:off + w], t, weight[off:off + w], mean_scale[off:off + w], st, ACT_NONE,
                                None, 0.0, None, None, dx, dw[off:off + w], db[off:off + w], da[off:off + w],
                                _gn_workspace(n, w, t.device))
            dxs.append(dx)
            off += w
        return (dw, db, da, None, *dxs)


def graph_norm_cat(xs: Sequence[torch.Tensor], weight, bias, mean_scale, eps: float = 1e-5):
    return _GraphNormCat.apply(weight, bias, mean_scale, eps, *xs)


_ids_checked = {}


def _validate_ids(ids: torch.Tensor, rows: int, what: str) -> None:
    """torch's nn.Embedding / index ops raise on an out-of-range index; the kernels cannot (no host sync), so
    the range is checked on the host once per index tensor (one sync, cached by address and version -- the
    node-id tensor of a dataset never changes) and skipped while a CUDA graph is being captured."""
    if ids.numel() == 0 or torch.cuda.is_current_stream_capturing():
        return
    import weakref
    base = ids._base if ids._base is not None else ids
    key = (ids.data_ptr(), ids._version, ids.numel(), str(ids.device), rows)
    hit = _ids_checked.get(key)
    if hit is not None and hit() is base:       # the same live tensor (not another one at a recycled address)
        return
    lo, hi = int(ids.min()), int(ids.max())
    if lo < 0 or hi >= rows:
        raise IndexError(f"{what}: index out of range [0, {rows}) (min {lo}, max {hi})")
    if len(_ids_checked) > 64:
        _ids_checked.clear()
    _ids_checked[key] = weakref.ref(base)


class _SpMMGraphNorm(torch.autograd.Function):
    """dropout(act(GraphNorm(adj @ x))) (impl/models.py:164-166) with the column statistics taken in the SpMM
    epilogue: no separate pass over `adj @ x` to find its mean / variance."""

    @staticmethod
    def forward(ctx, x, adj: CSRAdj, weight, bias, mean_scale, eps, act, p, training):
        x, _ = _rowmajor(_req(x, torch.float32, "x", 2))
        if x.shape[0] != adj.n:
            raise RuntimeError(f"spmm: x has {x.shape[0]} rows but the adjacency is {adj.n} x {adj.n}")
        weight, bias = _req(weight, torch.float32, "weight", 1), _req(bias, torch.float32, "bias", 1)
        mean_scale = _req(mean_scale, torch.float32, "mean_scale", 1)
        n, c = adj.n, x.shape[1]
        y = torch.empty((n, c), dtype=torch.float32, device=x.device)
        partial = _stats_table(c, x.device)
        nblk = _run_spmm(adj.rowptr, adj.col, adj.val, adj.plan, x, y, partial)
        drop_p, keep, rng, bits = _dropout_source(n, c, p, training, x.device)
        stats = torch.empty((6, c), dtype=torch.float32, device=x.device)
        _ops.graphnorm_stats_(partial, nblk, n, weight, bias, mean_scale, float(eps), keep, drop_p, rng, bits, stats)
        out = torch.empty((n, c), dtype=torch.float32, device=x.device)
        _ops.graphnorm_apply_(y, stats, act, keep, drop_p, bits, out)
        ctx.save_for_backward(y, weight, mean_scale, stats, keep, rng, bits)
        ctx.adj, ctx.cfg = adj, (act, drop_p)
        return out

    @staticmethod
    def backward(ctx, dout):
        y, weight, mean_scale, stats, keep, rng, bits = ctx.saved_tensors
        act, drop_p = ctx.cfg
        adj = ctx.adj
        dout, _ = _rowmajor(dout)
        n, c = y.shape
        dy = torch.empty((n, c), dtype=torch.float32, device=y.device)
        dw, db, da = torch.empty_like(weight), torch.empty_like(weight), torch.empty_like(weight)
        _ops.graphnorm_bwd_(dout, y, weight, mean_scale, stats, act, keep, drop_p, rng, bits, dy, dw, db, da,
                            _gn_workspace(n, c, y.device))
        gx = torch.empty_like(dy)
        _run_spmm(adj.rowptr_t, adj.col_t, adj.val_t, adj.plan_t, dy, gx)
        return gx, None, dw, db, da, None, None, None, None


def spmm_graph_norm(adj: CSRAdj, x, weight, bias, mean_scale, eps: float = 1e-5, act: int = ACT_NONE, p: float = 0.0,
                    training: bool = False):
    # the statistics-in-the-epilogue form pays where GraphNorm would otherwise take three launches; tiny matrices
    # (cluster kernel) and large power-of-two widths (cooperative kernel) are already one launch
    if adj.n == 0 or x.shape[0] == 0 or _lib.load().glass_graphnorm_launches(adj.n, x.shape[1]) == 1:
        return graph_norm(spmm(adj, x), weight, bias, mean_scale, eps, act, p, training)
    return _SpMMGraphNorm.apply(x, adj, weight, bias, mean_scale, eps, act, p, training)


def conv_fusable(k_in: int, h: int, path: Optional[int] = None) -> bool:
    """True when one GLASSConv layer (in width k_in, out width h) can run as the single autograd node below (the
    tcgen05 kernels incl. the dW kernel: needed for the accumulate-into-dX epilogue and the normalised loaders)."""
    path = _gemm_path if path is None else path
    if path == GEMM_SIMT:
        return False
    lib = _lib.load()
    return bool(lib.glass_pair_norm_operand_supported(k_in, 0, h)) and bool(lib.glass_pair_norm_operand_supported(h, k_in, h))


class _GlassConv(torch.autograd.Function):
    """One GLASSConv layer (impl/models.py:153-174) as ONE autograd node.  Default (fuse_norm False): pair GEMM, SpMM,
    one-launch GraphNorm + dropout, pair GEMM forward; backward = comb dX / dW, one-launch GraphNorm backward, A^T
    SpMM, trans dX / dW where the trans dX epilogue ACCUMULATES into the x_ gradient the comb GEMM wrote (x_ feeds both
    GEMMs; no separate add pass).  With fuse_norm (GLASS_B200_CONV_FUSED=1) the layer is FOUR kernels forward:
        pair GEMM (trans_fns + activation + label mix)                                         :158-162
        SpMM whose epilogue also emits the GraphNorm column sums                               :164-165
        finalize (normalisation constants + this call's dropout bits)
        pair GEMM over the virtual concat [dropout(GraphNorm(adj @ x)) | x_] whose loader applies the norm
        and the keep bits to the first operand on the fly (neither the normalised matrix nor the concat
        exists in memory)                                                                      :165-173
    and backward: comb dX / dW (the dW loader re-creates the normalised operand), GraphNorm backward, A^T SpMM,
    trans dX (ACCUMULATING into the gradient of x_, which fed both GEMMs) / dW."""

    @staticmethod
    def forward(ctx, x_, adj, tw0, tb0, tw1, tb1, gw, gb, gms, cw0, cb0, cw1, cb1, mask, z_ratio, act, eps, p,
                training, path, fuse_norm):
        x_, _ = _rowmajor(_req(x_, torch.float32, "x_", 2))
        mask = _req(mask, torch.uint8, "mask", 1)
        n, k_in = x_.shape
        h = tw0.shape[0]
        dev = x_.device
        if n != adj.n or mask.shape[0] != n:
            raise RuntimeError(f"GLASSConv: x_ has {n} rows, adjacency {adj.n}, mask {mask.shape[0]}")
        f32 = dict(dtype=torch.float32, device=dev)
        need_grad = any(ctx.needs_input_grad)
        # :158-162
        xm = torch.empty((n, h), **f32)
        acts = torch.empty((n, 2 * h), **f32) if (need_grad and act != ACT_NONE) else None
        _ops.pair_linear_mix_fwd_(x_, None, tw0, tb0, tw1, tb1, mask, float(z_ratio), act, path, xm, acts)
        y = torch.empty((n, h), **f32)
        drop_p, keep, rng, bits = _dropout_source(n, h, p, training, dev)
        stats = torch.empty((6, h), **f32)
        out = torch.empty((n, h), **f32)
        g = None
        if fuse_norm:
            # :164 with the statistics of :165 in the SpMM epilogue; :165-173 with the norm applied by the GEMM loader
            partial = _stats_table(h, dev)
            nblk = _run_spmm(adj.rowptr, adj.col, adj.val, adj.plan, xm, y, partial)
            _ops.graphnorm_stats_(partial, nblk, n, gw, gb, gms, float(eps), keep, drop_p, rng, bits, stats)
            _ops.pair_linear_mix_fwd_ex_(y, x_, cw0, cb0, cw1, cb1, mask, float(z_ratio), ACT_NONE, path, out, None,
                                         stats, bits, drop_p, ACT_NONE, None, None, 0.0, ACT_NONE)
        else:
            # :164, then GraphNorm + dropout as ONE launch that writes the normalised matrix (measured faster than
            # the loader fusion: the tcgen05 kernels are bound by their CUDA-core warps)
            _run_spmm(adj.rowptr, adj.col, adj.val, adj.plan, xm, y)
            g = torch.empty((n, h), **f32)
            _ops.graphnorm_fwd_(y, gw, gb, gms, float(eps), ACT_NONE, keep, drop_p, rng, bits, g, stats,
                                _gn_workspace(n, h, dev))
            _ops.pair_linear_mix_fwd_(g, x_, cw0, cb0, cw1, cb1, mask, float(z_ratio), ACT_NONE, path, out, None)
        ctx.save_for_backward(x_, acts, y, stats, keep, rng, bits, mask, tw0, tw1, gw, gms, cw0, cw1, g)
        ctx.adj, ctx.cfg = adj, (float(z_ratio), act, drop_p, path, fuse_norm)
        return out

    @staticmethod
    def backward(ctx, dout):
        x_, acts, y, stats, keep, rng, bits, mask, tw0, tw1, gw, gms, cw0, cw1, g = ctx.saved_tensors
        z_ratio, act, drop_p, path, fuse_norm = ctx.cfg
        adj = ctx.adj
        dout, _ = _rowmajor(dout)
        n, k_in = x_.shape
        h = tw0.shape[0]
        dev = dout.device
        f32 = dict(dtype=torch.float32, device=dev)
        lib = _lib.load()
        want_dx = ctx.needs_input_grad[0]
        # comb GEMM: gradient w.r.t. the normalised operand (dg) and the x_ half of the concat
        dg = torch.empty((n, h), **f32)
        dx_ = torch.empty((n, k_in), **f32) if want_dx else None
        dcw0, dcw1 = torch.empty_like(cw0), torch.empty_like(cw1)
        dcb0, dcb1 = torch.empty(h, **f32), torch.empty(h, **f32)
        ws = torch.empty(lib.glass_pair_linear_mix_bwd_workspace_bytes(n, h, h + k_in), dtype=torch.uint8, device=dev)
        if fuse_norm:
            _ops.pair_linear_mix_bwd_ex_(dout, None, y, x_, cw0, cw1, mask, z_ratio, ACT_NONE, path, dg, dx_, dcw0, dcb0,
                                         dcw1, dcb1, ws, stats, bits, drop_p, ACT_NONE, None, None, 0.0, ACT_NONE, 0, 0)
        else:
            _ops.pair_linear_mix_bwd_(dout, None, g, x_, cw0, cw1, mask, z_ratio, ACT_NONE, path, dg, dx_, dcw0, dcb0,
                                      dcw1, dcb1, ws)
        # GraphNorm + dropout backward, then A^T
        dy = torch.empty((n, h), **f32)
        dgw, dgb, dgms = torch.empty_like(gw), torch.empty_like(gw), torch.empty_like(gw)
        _ops.graphnorm_bwd_(dg, y, gw, gms, stats, ACT_NONE, keep, drop_p, rng, bits, dy, dgw, dgb, dgms,
                            _gn_workspace(n, h, dev))
        dxm = dg                                                   # dg is dead: reuse its storage
        _run_spmm(adj.rowptr_t, adj.col_t, adj.val_t, adj.plan_t, dy, dxm)
        # trans GEMM: its dX is ADDED to the x_ gradient the comb GEMM wrote above
        dtw0, dtw1 = torch.empty_like(tw0), torch.empty_like(tw1)
        dtb0, dtb1 = torch.empty(h, **f32), torch.empty(h, **f32)
        ws = torch.empty(lib.glass_pair_linear_mix_bwd_workspace_bytes(n, h, k_in), dtype=torch.uint8, device=dev)
        _ops.pair_linear_mix_bwd_ex_(dxm, acts, x_, None, tw0, tw1, mask, z_ratio, act, path, dx_, None, dtw0, dtb0,
                                     dtw1, dtb1, ws, None, None, 0.0, ACT_NONE, None, None, 0.0, ACT_NONE,
                                     1 if want_dx else 0, 0)
        return (dx_, None, dtw0, dtb0, dtw1, dtb1, dgw, dgb, dgms, dcw0, dcb0, dcw1, dcb1, None, None, None, None,
                None, None, None, None)


def glass_conv(x_, adj: CSRAdj, trans, gn, comb, mask, z_ratio: float, act: int, p: float, training: bool,
               path: Optional[int] = None):
    """Fused GLASSConv.forward; trans / comb = (w0, b0, w1, b1), gn = (weight, bias, mean_scale, eps)."""
    gw, gb, gms, eps = gn
    return _GlassConv.apply(x_, adj, *trans, gw, gb, gms, *comb, mask, z_ratio, act, eps, p, training,
                            _gemm_path if path is None else path, _conv_fused)


def pair_linear_mix_into(a1, a2, w0, b0, w1, b1, mask, z_ratio: float, act: int, out, acts=None, path=None):
    """Forward only, into caller-provided buffers (`acts` [n, 2h] receives the post-activation p0 | p1)."""
    _ops.pair_linear_mix_fwd_(_req(a1, torch.float32, "a1", 2), a2, w0, b0, w1, b1, _req(mask, torch.uint8, "mask", 1),
                              float(z_ratio), act, _gemm_path if path is None else path, out, acts)
    return out


def glass_conv_from_base(adj: CSRAdj, x_, y_u, delta, mask, gn, comb, z_ratio: float, path: Optional[int] = None):
    """Evaluation-mode GLASSConv for one label batch from the shared base (impl/models.py:164-173 with
    adj @ x = adj @ U + adj[:, labelled] @ delta[labelled]): correction SpMM with the statistics epilogue ->
    finalize -> comb GEMM with the norm applied on load (or a materialised norm on non-tcgen05 shapes)."""
    path = _gemm_path if path is None else path
    gw, gb, gms, eps = gn
    cw0, cb0, cw1, cb1 = comb
    n, h = y_u.shape
    dev = y_u.device
    y = torch.empty((n, h), dtype=torch.float32, device=dev)
    partial = _stats_table(h, dev)
    nblk = spmm_delta(adj, mask, delta, y_u, y, partial)
    stats = torch.empty((6, h), dtype=torch.float32, device=dev)
    _ops.graphnorm_stats_(partial, nblk, n, gw, gb, gms, float(eps), None, 0.0, None, None, stats)
    out = torch.empty((n, cw0.shape[0]), dtype=torch.float32, device=dev)
    if conv_fusable(x_.shape[1], h, path):
        _ops.pair_linear_mix_fwd_ex_(y, x_, cw0, cb0, cw1, cb1, mask, float(z_ratio), ACT_NONE, path, out, None,
                                     stats, None, 0.0, ACT_NONE, None, None, 0.0, ACT_NONE)
    else:
        g = torch.empty_like(y)
        _ops.graphnorm_apply_(y, stats, ACT_NONE, None, 0.0, None, g)
        _ops.pair_linear_mix_fwd_(g, x_, cw0, cb0, cw1, cb1, mask, float(z_ratio), ACT_NONE, path, out, None)
    return out


_embed_plans = {}


def _embed_plan(ids: torch.Tensor):
    """Ordered-accumulation plan of an id tensor (init path, cached by address / version): stable sort by id, runs of
    at most 128 sorted positions with one id, and for every distinct id its runs."""
    import weakref
    base = ids._base if ids._base is not None else ids          # the long-lived tensor the caller holds (x of the dataset)
    key = (ids.data_ptr(), ids._version, ids.numel(), str(ids.device))
    hit = _embed_plans.get(key)
    if hit is not None and hit[0]() is base:                     # same live tensor, not a recycled address
        return hit[1]
    n = ids.numel()
    sorted_ids, perm = torch.sort(ids, stable=True)
    pos = torch.arange(n, device=ids.device)
    change = torch.ones(n, dtype=torch.bool, device=ids.device)
    if n > 1:
        change[1:] = sorted_ids[1:] != sorted_ids[:-1]
    seg_start = torch.cummax(torch.where(change, pos, torch.zeros_like(pos)), 0).values
    run_start = change | ((pos - seg_start) % 128 == 0)
    run_begin = torch.nonzero(run_start).flatten()
    run_end = torch.cat((run_begin[1:], torch.tensor([n], device=ids.device)))
    run_id = sorted_ids[run_begin]
    uid, counts = torch.unique_consecutive(run_id, return_counts=True)
    first = torch.cumsum(counts, 0) - counts
    plan = dict(perm=perm.contiguous(), run_begin=run_begin.to(torch.int32), run_end=run_end.to(torch.int32),
                uid=uid.contiguous(), first=first.to(torch.int32), runs=counts.to(torch.int32))
    if len(_embed_plans) > 16:
        _embed_plans.clear()
    _embed_plans[key] = (weakref.ref(base), plan)
    return plan


class _Embedding(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, table):
        ids = _req(ids, torch.int64, "ids").reshape(-1)
        table = _req(table, torch.float32, "table", 2)
        _validate_ids(ids, table.shape[0], "embedding")
        out = torch.empty((ids.numel(), table.shape[1]), dtype=torch.float32, device=table.device)
        _ops.embedding_fwd_(table, ids, out)
        ctx.save_for_backward(ids)
        ctx.rows = table.shape[0]
        ctx.plan = None if torch.cuda.is_current_stream_capturing() else _embed_plan(ids)   # built outside any capture
        return out

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        dout, _ = _rowmajor(dout)
        h = dout.shape[1]
        dtable = torch.zeros((ctx.rows, h), dtype=torch.float32, device=dout.device)
        plan = ctx.plan
        if plan is None:                  # first use happened under capture: atomics (not bit-reproducible)
            _ops.embedding_bwd_(dout, ids, dtable)
            return None, dtable
        lib = _lib.load()
        run_sum = torch.empty((plan["run_begin"].numel(), h), dtype=torch.float32, device=dout.device)
        check(lib.glass_embedding_bwd_ordered(_p(dout), dout.stride(0), _p(plan["perm"]), _p(plan["run_begin"]),
                                              _p(plan["run_end"]), plan["run_begin"].numel(), _p(plan["uid"]),
                                              _p(plan["first"]), _p(plan["runs"]), plan["uid"].numel(), _p(run_sum),
                                              _p(dtable), ctx.rows, h, _stream()), "embedding_bwd_ordered")
        _count(2)
        return None, dtable


def embedding(ids: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    """table[ids] (nn.Embedding lookup of impl/models.py:248; also the emb[pos] gather of :348)."""
    return _Embedding.apply(ids, table)


class _SegmentPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, pos, mode):
        emb, _ = _rowmajor(_req(emb, torch.float32, "emb", 2))
        pos = _req(pos, torch.int64, "subG_node", 2)
        b, d = pos.shape[0], emb.shape[1]
        out = torch.empty((b, d), dtype=torch.float32, device=emb.device)
        cnt = torch.empty(b, dtype=torch.float32, device=emb.device)
        argmax = torch.empty((b, d), dtype=torch.int32, device=emb.device) if mode == POOL["max"] else None
        _ops.segment_pool_fwd_(emb, pos, mode, out, cnt, argmax)
        ctx.save_for_backward(pos, cnt, argmax)
        ctx.cfg = (mode, emb.shape[0])
        return out

    @staticmethod
    def backward(ctx, dout):
        pos, cnt, argmax = ctx.saved_tensors
        mode, n = ctx.cfg
        dout, _ = _rowmajor(dout)
        # node-centric ordered kernel: every row of demb is written (no zero-fill), additions in subgraph order
        demb = torch.empty((n, dout.shape[1]), dtype=torch.float32, device=dout.device)
        mark = torch.empty(_lib.load().glass_segment_pool_bwd_scratch_bytes(pos.shape[0], n), dtype=torch.uint8,
                           device=dout.device)
        _ops.segment_pool_bwd_(dout, pos, mode, cnt, argmax, demb, mark)
        return demb, None, None


def segment_pool(emb: torch.Tensor, subG_node: torch.Tensor, mode: str) -> torch.Tensor:
    """GLASS.Pool (impl/models.py:346-350) fused: pad2batch + emb[pos] + {Add,Mean,Max,Size}Pool."""
    if mode not in POOL:
        raise NotImplementedError(mode)  # GLASSTest.py:171
    return _SegmentPool.apply(emb, subG_node, POOL[mode])


class _SegmentPoolBatch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, batch, n_seg, mode):
        x, _ = _rowmajor(_req(x, torch.float32, "x", 2))
        batch = _req(batch, torch.int64, "batch", 1)
        d = x.shape[1]
        out = torch.empty((n_seg, d), dtype=torch.float32, device=x.device)
        cnt = torch.empty(n_seg, dtype=torch.float32, device=x.device)
        argmax = torch.empty((n_seg, d), dtype=torch.int32, device=x.device) if mode == POOL["max"] else None
        _ops.segment_pool_batch_fwd_(x, batch, n_seg, mode, out, cnt, argmax)
        ctx.save_for_backward(batch, cnt, argmax)
        ctx.cfg = (mode, n_seg, x.shape[0])
        return out

    @staticmethod
    def backward(ctx, dout):
        batch, cnt, argmax = ctx.saved_tensors
        mode, n_seg, m = ctx.cfg
        dout, _ = _rowmajor(dout)
        dx = torch.empty((m, dout.shape[1]), dtype=torch.float32, device=dout.device)
        _ops.segment_pool_batch_bwd_(dout, batch, n_seg, mode, cnt, argmax, dx)
        return dx, None, None, None


NORM_POOL_MODES = ("sum", "mean", "size")


class _NormPoolCat(torch.autograd.Function):
    """pool(GraphNorm(cat(xs))[subG_node]) for a GraphNorm whose output feeds only the pooling (the model's last one,
    impl/models.py:266 / :272 -> :346-350): statistics over all rows, normalisation only on the gathered rows; backward
    forms the norm's column sums from the pooled gradients and writes dx in one pass (csrc/pool.cu).  xs are the COLUMN
    BLOCKS of a virtual concat (JK, :263-267; one block otherwise): GraphNorm is per column, so block l is normalised
    with the parameter slice [off_l, off_l + w_l) and pooled into the same columns of the output.  One autograd node:
    no parameter slicing / gradient re-assembly kernels."""

    @staticmethod
    def forward(ctx, weight, bias, mean_scale, eps, pos, mode, *xs):
        pos = _req(pos, torch.int64, "subG_node", 2)
        weight, bias = _req(weight, torch.float32, "weight", 1), _req(bias, torch.float32, "bias", 1)
        mean_scale = _req(mean_scale, torch.float32, "mean_scale", 1)
        xs = [_rowmajor(_req(x, torch.float32, "x", 2))[0] for x in xs]
        n, dev = xs[0].shape[0], xs[0].device
        d_total = sum(x.shape[1] for x in xs)
        lib = _lib.load()
        b = pos.shape[0]
        out = torch.empty((b, d_total), dtype=torch.float32, device=dev)
        cnt = torch.empty(b, dtype=torch.float32, device=dev)
        saved, off = [], 0
        for x in xs:
            c = x.shape[1]
            partial = torch.empty((2 * c, lib.glass_graphnorm_partials_ld()), dtype=torch.float64, device=dev)
            nblk = graphnorm_partials(x, partial)
            stats = torch.empty((6, c), dtype=torch.float32, device=dev)
            _ops.graphnorm_stats_(partial, nblk, n, weight[off:off + c], bias[off:off + c], mean_scale[off:off + c],
                                  float(eps), None, 0.0, None, None, stats)
            ysum = torch.empty((b, c), dtype=torch.float32, device=dev)
            _ops.norm_pool_fwd_(x, stats, pos, mode, out[:, off:off + c], cnt, ysum)
            saved += [x, stats, ysum]
            off += c
        ctx.save_for_backward(weight, mean_scale, pos, cnt, *saved)
        ctx.mode, ctx.n_blocks = mode, len(xs)
        return out

    @staticmethod
    def backward(ctx, dout):
        weight, mean_scale, pos, cnt = ctx.saved_tensors[:4]
        rest = ctx.saved_tensors[4:]
        dout, _ = _rowmajor(dout)
        dw, db, dms = torch.empty_like(weight), torch.empty_like(weight), torch.empty_like(weight)
        lib = _lib.load()
        dxs, off = [], 0
        for i in range(ctx.n_blocks):
            x, stats, ysum = rest[3 * i:3 * i + 3]
            n, c = x.shape
            dx = torch.empty((n, c), dtype=torch.float32, device=x.device)
            scratch = torch.empty(lib.glass_segment_pool_bwd_scratch_bytes(pos.shape[0], n), dtype=torch.uint8, device=x.device)
            _ops.norm_pool_bwd_(dout[:, off:off + c], pos, ctx.mode, cnt, ysum, x, stats, weight[off:off + c],
                                mean_scale[off:off + c], dx, dw[off:off + c], db[off:off + c], dms[off:off + c], scratch)
            dxs.append(dx)
            off += c
        return (dw, db, dms, None, None, None, *dxs)


def graph_norm_pool_cat(xs, weight, bias, mean_scale, eps: float, pos: torch.Tensor, mode: str) -> torch.Tensor:
    """segment_pool(graph_norm(cat(xs, -1), ...), pos, mode) without the concat, the normalised matrix or the pooled
    gradient matrix; mode in NORM_POOL_MODES, every block at most 256 columns wide."""
    if mode not in NORM_POOL_MODES:
        raise NotImplementedError(mode)
    xs = list(xs)
    if xs[0].shape[0] == 0 or max(x.shape[1] for x in xs) > 256:
        return segment_pool(graph_norm_cat(xs, weight, bias, mean_scale, eps), pos, mode)
    return _NormPoolCat.apply(weight, bias, mean_scale, eps, pos, POOL[mode], *xs)


def graph_norm_pool(x, weight, bias, mean_scale, eps: float, pos: torch.Tensor, mode: str) -> torch.Tensor:
    """segment_pool(graph_norm(x, weight, bias, mean_scale, eps), pos, mode) as one operator (the single-block case of
    graph_norm_pool_cat); mode in NORM_POOL_MODES, at most 256 columns."""
    return graph_norm_pool_cat([x], weight, bias, mean_scale, eps, pos, mode)


def segment_pool_batch(x: torch.Tensor, batch: torch.Tensor, mode: str, size: Optional[int] = None):
    """PoolModule.forward(x, batch) (impl/models.py:287-292) for rows already gathered; batch sorted ascending.
    Like PyG, the number of segments is batch.max()+1 unless `size` is given (one host sync)."""
    if mode not in POOL:
        raise NotImplementedError(mode)
    n_seg = int(batch.max().item()) + 1 if size is None else int(size)
    return _SegmentPoolBatch.apply(x, batch, n_seg, POOL[mode])


# ---------------------------------------------------------------------------------------------
# labels / index helpers (no gradients)
# ---------------------------------------------------------------------------------------------
def maxzoz(n_node: int, pos: torch.Tensor, with_mask: bool = False):
    """Max-zero-one labels (impl/utils.py:32-45): int64 z[N]; optionally also the uint8 mask z > 0."""
    pos = _req(pos, torch.int64, "pos")
    z = torch.empty(n_node, dtype=torch.int64, device=pos.device)
    mask = torch.empty(n_node, dtype=torch.uint8, device=pos.device) if with_mask else None
    _ops.maxzoz_(pos, z, mask)
    return (z, mask) if with_mask else z


def label_mask(z: torch.Tensor) -> torch.Tensor:
    """uint8 (z > 0.5) of impl/models.py:246."""
    z = _req(z, torch.int64, "z").reshape(-1)
    mask = torch.empty(z.numel(), dtype=torch.uint8, device=z.device)
    _ops.label_mask_(z, mask)
    return mask


def pad2batch(pad: torch.Tensor):
    """impl/utils.py:18-29 on the GPU (one host sync for the output length, as in the reference)."""
    pad = _req(pad, torch.int64, "pad", 2)
    total = pad.numel()
    batch = torch.empty(total, dtype=torch.int64, device=pad.device)
    pos = torch.empty(total, dtype=torch.int64, device=pad.device)
    n_valid = torch.zeros(1, dtype=torch.int64, device=pad.device)
    _ops.pad2batch_(pad, batch, pos, n_valid)
    m = int(n_valid.item())
    return batch[:m], pos[:m]
