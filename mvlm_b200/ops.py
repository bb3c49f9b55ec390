"""Thin Python wrappers over single C-ABI stage calls (used by tests and the host classes).

Tensors are torch CUDA tensors used purely as device buffers.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ConvArgs, check, cur_stream, ptr


def pack_conv_weight(w_oihw: torch.Tensor, cout_pad: int, cin_pad: int) -> torch.Tensor:
    """fp32 OIHW (cuda) -> bf16 [cout_pad, kw, kh, cin_pad] (cuda)."""
    lib = _lib.load()
    w = w_oihw.contiguous().float()
    cout, cin, kh, kw = w.shape
    out = torch.empty((cout_pad, kw, kh, cin_pad), dtype=torch.bfloat16, device=w.device)
    check(lib.mvlm_pack_conv_weight(ptr(w), cout, cin, kh, kw, cout_pad, cin_pad, ptr(out), cur_stream()),
          "mvlm_pack_conv_weight")
    return out


def conv2d_bf16(x: torch.Tensor, wpacked: torch.Tensor, *, cin: int | None = None, n_tile: int,
                kh: int = 3, kw: int = 3, y_off0: int | None = None, x_off0: int | None = None,
                bias=None, pre=None, res1=None, res2=None, out_raw=None, post=None,
                out_f32=None, argmax_keys=None, cout_real: int | None = None,
                up=(1, 1, 0, 0)) -> None:
    """One fused conv launch.

    x          : (N,H,W,Cs) bf16; the first `cin` channels are read.
    pre / post : (scale f32[cout_pad], shift f32[cout_pad], out bf16 (N,H,W,Cs'), channel offset)
    res1/res2  : (tensor bf16 (N,H,W,Cs'), channel offset)
    out_raw    : (tensor bf16 (N,H,W,Cs'), channel offset)
    """
    lib = _lib.load()
    n, h, w, cs = x.shape
    a = ConvArgs()
    a.in_ = ptr(x)
    a.n, a.h, a.w, a.cin, a.in_cs = n, h, w, (cin if cin is not None else cs), cs
    a.wpacked = ptr(wpacked)
    a.cout_pad = wpacked.shape[0]
    a.n_tile, a.kh, a.kw = n_tile, kh, kw
    a.y_off0 = -(kh // 2) if y_off0 is None else y_off0
    a.x_off0 = -(kw // 2) if x_off0 is None else x_off0
    a.bias = ptr(bias)
    if pre is not None:
        a.pre_scale, a.pre_shift, a.out_pre = ptr(pre[0]), ptr(pre[1]), ptr(pre[2])
        a.pre_cs, a.pre_co = pre[2].shape[-1], pre[3]
    if res1 is not None:
        a.res1, a.res1_cs, a.res1_co = ptr(res1[0]), res1[0].shape[-1], res1[1]
    if res2 is not None:
        a.res2, a.res2_cs, a.res2_co = ptr(res2[0]), res2[0].shape[-1], res2[1]
    if out_raw is not None:
        a.out_raw, a.raw_cs, a.raw_co = ptr(out_raw[0]), out_raw[0].shape[-1], out_raw[1]
    if post is not None:
        a.post_scale, a.post_shift, a.out_post = ptr(post[0]), ptr(post[1]), ptr(post[2])
        a.post_cs, a.post_co = post[2].shape[-1], post[3]
    a.out_f32 = ptr(out_f32)
    a.argmax_keys = ptr(argmax_keys)
    a.cout_real = cout_real if cout_real is not None else wpacked.shape[0]
    a.up_sy, a.up_sx, a.up_py, a.up_px = up
    check(lib.mvlm_conv2d_bf16(C.byref(a), cur_stream()), "mvlm_conv2d_bf16")
