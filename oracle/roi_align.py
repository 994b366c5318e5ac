"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/roi_align_oracle.c (built with gcc on demand)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "libroi_align_oracle.so")
_SRC = os.path.join(_HERE, "roi_align_oracle.c")
_lib = None


def build():
    os.makedirs(_BUILD, exist_ok=True)
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        _lib.roi_align_forward_oracle.argtypes = [fp, fp] + [ctypes.c_int] * 6 + [ctypes.c_float, ctypes.c_int, fp]
        _lib.roi_align_backward_oracle.argtypes = [fp, fp] + [ctypes.c_int] * 7 + [ctypes.c_float, ctypes.c_int, fp]
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def roi_align_forward(inp, rois, spatial_scale, ph, pw, sampling_ratio):
    """inp [N,C,H,W] float32 ndarray, rois [K,5] -> [K,C,ph,pw]"""
    inp = np.ascontiguousarray(inp, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    N, C, H, W = inp.shape
    K = rois.shape[0]
    out = np.empty((K, C, ph, pw), dtype=np.float32)
    _load().roi_align_forward_oracle(_p(inp), _p(rois), K, C, H, W, ph, pw, float(spatial_scale), int(sampling_ratio), _p(out))
    return out


def roi_align_backward(grad, rois, spatial_scale, ph, pw, N, C, H, W, sampling_ratio):
    grad = np.ascontiguousarray(grad, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    K = rois.shape[0]
    gin = np.empty((N, C, H, W), dtype=np.float32)
    _load().roi_align_backward_oracle(_p(grad), _p(rois), K, N, C, H, W, ph, pw, float(spatial_scale), int(sampling_ratio), _p(gin))
    return gin
