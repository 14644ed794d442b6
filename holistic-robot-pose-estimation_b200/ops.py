"""Operator-level Python wrappers over the C ABI (used by the tests and by the operator-level shims).

Tensors are torch CUDA tensors; only their data_ptr() crosses the boundary.  Activations are bf16 NHWC with the
channel count padded to 16 / 32 / a multiple of 64 (`pad_channels`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import ConvDesc, ConvEpilogue, check

CONV, DECONV_K4S2P1, STEM_S2D, CONV_UP2 = 0, 1, 2, 3
IMPL_TCGEN05, IMPL_SIMT_CHECK = 0, 1


def pad_channels(c: int) -> int:
    if c <= 16:
        return 16
    if c <= 32:
        return 32
    return (c + 63) // 64 * 64


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def pack_conv_weights(desc: ConvDesc, weight: torch.Tensor, cin_ref: int) -> torch.Tensor:
    """weight: fp32 CPU tensor in the reference layout -> packed bf16 (as int16 bits) CPU tensor."""
    L = _lib.lib()
    n = C.c_int64(0)
    check(L.hrp_conv_packed_weight_elems(C.byref(desc), C.byref(n)))
    w = np.ascontiguousarray(weight.detach().cpu().float().numpy())
    out = np.zeros(n.value, dtype=np.uint16)
    check(L.hrp_conv_pack_weights(C.byref(desc), C.c_int32(cin_ref),
                                  w.ctypes.data_as(C.POINTER(C.c_float)),
                                  out.ctypes.data_as(C.POINTER(C.c_uint16))))
    return torch.from_numpy(out.view(np.int16)).view(torch.bfloat16)


class ConvOp:
    """One planned convolution: y = epilogue(conv(x)).  See include/hrp.h (hrp_conv_*)."""

    def __init__(self, x: torch.Tensor, weight: torch.Tensor, *, kind: int = CONV, stride: int = 1, pad: int = 0,
                 relu: bool = False, scale: torch.Tensor | None = None, bias: torch.Tensor | None = None,
                 pre=(), up=(), post: torch.Tensor | None = None, pool: bool = False, write_out: bool = True,
                 kernel_hw: tuple[int, int] | None = None):
        assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4
        L = _lib.lib()
        B, Hin, Win, Cin = x.shape
        if kind == DECONV_K4S2P1:
            cin_ref, Cout = weight.shape[0], weight.shape[1]
            kh = kw = 4
        else:
            Cout, cin_ref = weight.shape[0], weight.shape[1]
            kh, kw = (weight.shape[2], weight.shape[3]) if kernel_hw is None else kernel_hw
        self.desc = ConvDesc(kind, B, Hin, Win, Cin, Cout, kh, kw, stride, pad, int(relu))
        self.x = x
        self.w_packed = pack_conv_weights(self.desc, weight, cin_ref).cuda()
        dev = x.device
        self.scale = (torch.ones(Cout) if scale is None else scale).float().contiguous().to(dev)
        self.bias = (torch.zeros(Cout) if bias is None else bias).float().contiguous().to(dev)
        self.keep = [x, self.w_packed, self.scale, self.bias, *pre, *[u for u, _ in up], post]
        epi = ConvEpilogue()
        epi.scale = self.scale.data_ptr()
        epi.bias = self.bias.data_ptr()
        for i, t in enumerate(pre):
            epi.pre[i] = t.data_ptr()
        for i, (t, s) in enumerate(up):
            epi.up[i] = t.data_ptr()
            epi.up_shift[i] = s
        epi.post = post.data_ptr() if post is not None else None
        # output dims: create with a dummy out first to learn Hout/Wout is awkward -> compute here
        if kind in (DECONV_K4S2P1, CONV_UP2):
            Hout, Wout = 2 * Hin, 2 * Win
        elif kind == STEM_S2D:
            Hout, Wout = Hin, Win
        else:
            Hout = (Hin + 2 * pad - kh) // stride + 1
            Wout = (Win + 2 * pad - kw) // stride + 1
        self.out = torch.empty(B, Hout, Wout, Cout, dtype=torch.bfloat16, device=dev) if write_out else None
        self.pool_out = torch.zeros(B, Cout, dtype=torch.float32, device=dev) if pool else None
        epi.out = self.out.data_ptr() if self.out is not None else None
        epi.pool_out = self.pool_out.data_ptr() if self.pool_out is not None else None
        self.handle = C.c_void_p(0)
        with torch.cuda.device(dev):
            check(L.hrp_conv_create(C.byref(self.desc), _ptr(x), _ptr(self.w_packed), C.byref(epi),
                                    C.byref(self.handle)))

    def run(self, impl: int = IMPL_TCGEN05):
        if self.pool_out is not None:
            self.pool_out.zero_()
        with torch.cuda.device(self.x.device):
            check(_lib.lib().hrp_conv_run(self.handle, C.c_int32(impl), _stream()))
        return self.out if self.out is not None else self.pool_out

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().hrp_conv_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def pack_input_s2d(x: torch.Tensor) -> torch.Tensor:
    """(B,3,H,W) fp32 NCHW -> (B,H/2,W/2,16) bf16 space-to-depth (include/hrp.h: hrp_pack_input_s2d)."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 3
    B, _, H, W = x.shape
    out = torch.empty(B, H // 2, W // 2, 16, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().hrp_pack_input_s2d(_ptr(x), _ptr(out), B, H, W, _stream()))
    return out


def maxpool3x3s2(x: torch.Tensor) -> torch.Tensor:
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous()
    B, H, W, Cc = x.shape
    out = torch.empty(B, H // 2, W // 2, Cc, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().hrp_maxpool3x3s2(_ptr(x), _ptr(out), B, H, W, Cc, _stream()))
    return out


def nchw_to_nhwc_bf16(x: torch.Tensor, cpad: int | None = None) -> torch.Tensor:
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    B, Cc, H, W = x.shape
    cpad = pad_channels(Cc) if cpad is None else cpad
    out = torch.empty(B, H, W, cpad, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().hrp_nchw_f32_to_nhwc_bf16(_ptr(x), _ptr(out), B, Cc, H, W, cpad, _stream()))
    return out


def nhwc_bf16_to_nchw(x: torch.Tensor, c: int | None = None) -> torch.Tensor:
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous()
    B, H, W, cpad = x.shape
    c = cpad if c is None else c
    out = torch.empty(B, c, H, W, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().hrp_nhwc_bf16_to_nchw_f32(_ptr(x), _ptr(out), B, c, H, W, cpad, _stream()))
    return out
