"""Operator-level entry point: the fused implicit-GEMM causal Conv1d over per-slot context
buffers (`conan_conv_gemm` of the C ABI), callable on torch CUDA tensors.  Used by the parity
tests to drive both engines (FFMA and tcgen05) on arbitrary shapes."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

ACT = {"none": 0, "relu": 1, "lrelu": 2, "gelu": 3, "tanh": 4}
ENGINE_FFMA, ENGINE_TC = 0, 1


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def conv_gemm(ctx: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], *, k: int, dil: int, L: int, row0: int,
              slot_ids: Optional[torch.Tensor] = None, n_streams: Optional[int] = None, engine: int = ENGINE_FFMA,
              scale: float = 1.0, act: str = "none", slope: float = 0.0, res: Optional[torch.Tensor] = None,
              rowmask: Optional[torch.Tensor] = None, out_scale: float = 1.0, y: Optional[torch.Tensor] = None,
              accumulate: bool = False, y2: Optional[torch.Tensor] = None, y2_row0: int = 0, act2: str = "none",
              slope2: float = 0.0, x_split: bool = False, acc_scale: float = 1.0, y2_split: bool = False, res_inv_slope: float = 0.0,
              res2: Optional[torch.Tensor] = None):
    """ctx [slots, rows, cin] (fp32 or fp16), w_packed [cout, k*cin] (same dtype), bias [cout] fp32.
    y [slots, L, cout] fp32 (optional), y2 [slots, rows2, cout] fp32/fp16 written at rows y2_row0.. (optional),
    res [slots, L, cout] fp32 (optional), rowmask [slots, L] fp32 (optional)."""
    lib = _lib.load()
    if x_split:          # ctx [2, slots, rows, cin] fp16 (hi plane, lo plane); w_packed [cout, 3*k*cin]
        assert ctx.dim() == 4 and ctx.shape[0] == 2 and ctx.dtype == torch.float16
        _, slots, rows, cin = ctx.shape
    else:
        slots, rows, cin = ctx.shape
    cout = w_packed.shape[0]
    assert ctx.is_cuda and ctx.is_contiguous() and w_packed.is_contiguous() and w_packed.shape[1] == (3 if x_split else 1) * k * cin
    assert ctx.dtype == w_packed.dtype and ctx.dtype in (torch.float32, torch.float16)
    p = _lib.ConvParams()
    p.x, p.x_slot_stride, p.x_row_stride, p.x_rows = _p(ctx), rows * cin, cin, rows
    p.x_is_half = int(ctx.dtype == torch.float16)
    p.row0, p.L, p.cin, p.k, p.dil, p.cout = row0, L, cin, k, dil, cout
    p.w, p.bias = _p(w_packed), _p(bias)
    p.n_streams = n_streams if n_streams is not None else (slot_ids.numel() if slot_ids is not None else slots)
    p.slot_ids, p.n_slots = _p(slot_ids), slots
    p.scale, p.act, p.slope = scale, ACT[act], slope
    if res is not None:
        assert res.dtype in (torch.float32, torch.float16) and res.shape == (slots, L, cout) and res.is_contiguous()
        p.res, p.res_slot_stride, p.res_row_stride = _p(res), L * cout, cout
        p.res_is_half, p.res_inv_slope = int(res.dtype == torch.float16), res_inv_slope
    if rowmask is not None:
        assert rowmask.dtype == torch.float32 and rowmask.shape == (slots, L)
        p.rowmask, p.mask_slot_stride = _p(rowmask), L
    p.out_scale = out_scale
    if res2 is not None:
        assert res2.dtype in (torch.float32, torch.float16) and res2.shape == (slots, L, cout) and res2.is_contiguous()
        p.res2, p.res2_slot_stride, p.res2_row_stride, p.res2_is_half = _p(res2), L * cout, cout, int(res2.dtype == torch.float16)
    if y is not None:
        assert y.dtype in (torch.float32, torch.float16) and y.shape == (slots, L, cout) and y.is_contiguous()
        p.y, p.y_slot_stride, p.y_row_stride, p.y_row0 = _p(y), L * cout, cout, 0
        p.y_is_half = int(y.dtype == torch.float16)
    p.accumulate = int(accumulate)
    if y2 is not None and y2_split:      # y2 [2, slots, rows2, cout] fp16
        assert y2.dim() == 4 and y2.shape[0] == 2 and y2.dtype == torch.float16 and y2.is_contiguous()
        p.y2, p.y2_slot_stride, p.y2_row_stride, p.y2_row0 = _p(y2), y2.shape[2] * cout, cout, y2_row0
        p.y2_is_half, p.y2_split, p.y2_lo_off = 1, 1, y2.shape[1] * y2.shape[2] * cout
    elif y2 is not None:
        assert y2.shape[0] == slots and y2.shape[2] == cout and y2.is_contiguous()
        p.y2, p.y2_slot_stride, p.y2_row_stride, p.y2_row0 = _p(y2), y2.shape[1] * cout, cout, y2_row0
        p.y2_is_half = int(y2.dtype == torch.float16)
    if x_split:
        p.x_split, p.x_lo_slot_off = 1, slots
    p.acc_scale = acc_scale
    p.act2, p.slope2 = ACT[act2], slope2
    st = C.c_void_p(torch.cuda.current_stream(ctx.device).cuda_stream)
    _lib.check(lib.conan_conv_gemm(C.byref(p), engine, st), "conv_gemm")
