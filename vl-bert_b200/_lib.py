"""ctypes binding of libvlbert_b200.so (the C ABI in include/vlbert_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a RuntimeError
is raised (the reference's native op raises RuntimeError through AT_ASSERTM / THCudaCheck,
common/lib/roi_pooling/cuda/ROIAlign_cuda.cu:262-264,297).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvlbert_b200.so")

_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_float = ctypes.c_float
c_int64 = ctypes.c_int64
c_uint32 = ctypes.c_uint32


def _declare(lib):
    lib.vlb_abi_version.restype = c_int
    lib.vlb_abi_version.argtypes = []
    lib.vlb_last_error_string.restype = ctypes.c_char_p
    lib.vlb_last_error_string.argtypes = []
    lib.vlb_launch_count.restype = c_int64
    lib.vlb_launch_count.argtypes = []
    lib.vlb_gemm_bf16.restype = c_int
    lib.vlb_gemm_bf16.argtypes = [c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                                  c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                  c_void_p, c_int, c_float, c_int, c_int, c_void_p]
    lib.vlb_debug_gemm_desc.restype = None
    lib.vlb_debug_gemm_desc.argtypes = [c_uint32, c_uint32, c_uint32]


def lib():
    """Load (once) and return the ctypes handle; raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "vlbert_b200: %s not found -- run `python vl-bert_b200/build.py` (there is no CPU "
                "fallback)" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        _declare(handle)
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().vlb_last_error_string().decode("utf-8", "replace")
        raise RuntimeError("vlbert_b200 error %d: %s" % (rc, msg))


def last_error():
    return lib().vlb_last_error_string().decode("utf-8", "replace")


def launch_count():
    return int(lib().vlb_launch_count())
