"""Thin torch-tensor wrappers over the U-Net C-ABI primitives (NHWC bf16).  Used by the tests and
by eld_b200.arch; all compute happens in libeld_b200.so."""
import ctypes

import torch

from . import _lib

PACK_CONV_FPROP, PACK_CONV_DGRAD, PACK_DECONV_FPROP, PACK_DECONV_DGRAD = 0, 1, 2, 3
ACT_NONE, ACT_LRELU, ACT_MASK = 0, 1, 2


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ctx(t):
    return _lib.ctx(t.device.index or 0)


def pack_weights(w, kind):
    """w: fp32 cuda, Conv2d OIHW [cout,cin,3,3] or ConvTranspose2d IOHW [cin,cout,2,2] -> bf16 operand."""
    w = w.contiguous().float()
    if kind in (PACK_CONV_FPROP, PACK_CONV_DGRAD):
        cout, cin = w.shape[0], w.shape[1]
        shape = (cout, 9 * cin) if kind == PACK_CONV_FPROP else (cin, 9 * cout)
    else:
        cin, cout = w.shape[0], w.shape[1]
        shape = (4 * cout, cin) if kind == PACK_DECONV_FPROP else (cin, 4 * cout)
    out = torch.empty(shape, dtype=torch.bfloat16, device=w.device)
    _lib.check(_lib.load().eld_pack_weights(_ctx(w), w.data_ptr(), out.data_ptr(), cout, cin, kind, _st()),
               'eld_pack_weights')
    return out


def conv3x3(x, x_c0, cin, w_packed, bias, y, y_c0, cout, act=ACT_NONE, aux=None, aux_c0=0):
    """x,y: NHWC bf16 buffers [n,h,w,pitch]; uses channels [c0, c0+c)."""
    n, h, w, xp = x.shape
    assert y.shape[:3] == x.shape[:3] and x.dtype == torch.bfloat16 and y.dtype == torch.bfloat16
    assert x.is_contiguous() and y.is_contiguous()
    rc = _lib.load().eld_conv3x3_bf16(_ctx(x), x.data_ptr(), xp, x_c0, cin, w_packed.data_ptr(),
                                      bias.data_ptr() if bias is not None else None,
                                      y.data_ptr(), y.shape[3], y_c0, cout, n, h, w, act,
                                      aux.data_ptr() if aux is not None else None,
                                      aux.shape[3] if aux is not None else 0, aux_c0, _st())
    _lib.check(rc, 'eld_conv3x3_bf16')
    return y


def deconv2x2(x, x_c0, cin, w_packed, bias, y, y_c0, cout):
    n, h, w, xp = x.shape
    assert y.shape[0] == n and y.shape[1] == 2 * h and y.shape[2] == 2 * w
    rc = _lib.load().eld_deconv2x2_bf16(_ctx(x), x.data_ptr(), xp, x_c0, cin, w_packed.data_ptr(),
                                        bias.data_ptr() if bias is not None else None,
                                        y.data_ptr(), y.shape[3], y_c0, cout, n, h, w, _st())
    _lib.check(rc, 'eld_deconv2x2_bf16')
    return y


def deconv2x2_dgrad(dy, dy_c0, cout, w_packed, dx, dx_c0, cin, act=ACT_NONE, aux=None, aux_c0=0):
    n, h, w, _ = dx.shape
    assert dy.shape[1] == 2 * h and dy.shape[2] == 2 * w
    rc = _lib.load().eld_deconv2x2_dgrad_bf16(_ctx(dy), dy.data_ptr(), dy.shape[3], dy_c0, cout,
                                              w_packed.data_ptr(), dx.data_ptr(), dx.shape[3], dx_c0, cin,
                                              n, h, w, act, aux.data_ptr() if aux is not None else None,
                                              aux.shape[3] if aux is not None else 0, aux_c0, _st())
    _lib.check(rc, 'eld_deconv2x2_dgrad_bf16')
    return dx


def conv3x3_wgrad(x, x_c0, cin, dz, dz_c0, cout, dw):
    """dw: f32 [cout,cin,3,3], accumulated into."""
    n, h, w, xp = x.shape
    rc = _lib.load().eld_conv3x3_wgrad_bf16(_ctx(x), x.data_ptr(), xp, x_c0, cin, dz.data_ptr(), dz.shape[3], dz_c0,
                                            cout, dw.data_ptr(), n, h, w, _st())
    _lib.check(rc, 'eld_conv3x3_wgrad_bf16')
    return dw


def deconv2x2_wgrad(x, x_c0, cin, dy, dy_c0, cout, dw):
    """x coarse [n,h,w,*], dy fine [n,2h,2w,*]; dw: f32 [cin,cout,2,2], accumulated into."""
    n, h, w, xp = x.shape
    rc = _lib.load().eld_deconv2x2_wgrad_bf16(_ctx(x), x.data_ptr(), xp, x_c0, cin, dy.data_ptr(), dy.shape[3], dy_c0,
                                              cout, dw.data_ptr(), n, h, w, _st())
    _lib.check(rc, 'eld_deconv2x2_wgrad_bf16')
    return dw
