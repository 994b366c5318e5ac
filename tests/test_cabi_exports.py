"""CPU: the C-ABI library builds, loads, and exports every symbol include/vlbert_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vlbert_b200.h")
LIB = os.path.join(ROOT, "vl-bert_b200", "libvlbert_b200.so")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vlb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import importlib.util
        spec = importlib.util.spec_from_file_location("vlb_build", os.path.join(ROOT, "vl-bert_b200", "build.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        m.build()
    return ctypes.CDLL(LIB)


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("vlb_gemm_bf16", "vlb_mhsa_forward", "vlb_mhsa_backward", "vlb_layernorm_forward", "vlb_layernorm_backward",
                 "vlb_pack_index", "vlb_pack_forward", "vlb_pack_backward", "vlb_gather_rows", "vlb_roi_align_forward",
                 "vlb_roi_align_backward", "vlb_region_operand", "vlb_bert_layer_forward", "vlb_bert_layer_backward",
                 "vlb_last_error_string"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing


def test_no_compute_needed_calls(lib):
    lib.vlb_abi_version.restype = ctypes.c_int
    assert lib.vlb_abi_version() == 2
    lib.vlb_last_error_string.restype = ctypes.c_char_p
    assert isinstance(lib.vlb_last_error_string(), bytes)
    lib.vlb_bert_layer_backward_workspace.restype = ctypes.c_int64
    lib.vlb_bert_layer_backward_workspace.argtypes = [ctypes.c_int] * 3
    M, H, I = 6464, 768, 3072
    assert lib.vlb_bert_layer_backward_workspace(M, H, I) >= M * (4 * H + I + 3 * H) * 2 + M * 3 * H * 4


def test_bad_arguments_return_error_codes_not_crashes(lib):
    lib.vlb_gemm_bf16.restype = ctypes.c_int
    P, I, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.vlb_gemm_bf16.argtypes = [I, I, I, I, P, I, P, I, P, I, I, P, P, I, I, I, P, I, F, I, I, P]
    rc = lib.vlb_gemm_bf16(7, 128, 128, 64, None, 64, None, 64, None, 128, 0, None, None, 0, 0, 0, None, 0, 1.0, 1, 0, None)
    assert rc == -1
    assert b"mode" in lib.vlb_last_error_string()
    rc = lib.vlb_gemm_bf16(0, 128, 130, 64, 16, 64, 16, 64, 16, 128, 0, None, None, 0, 0, 0, None, 0, 1.0, 1, 0, None)
    assert rc == -1  # N not a multiple of 8


def test_python_binding_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import vlbert_b200
    import vlbert_oracle as vo
    m = vlbert_b200.VisualLinguisticBert(vo.default_config(num_hidden_layers=1))
    ids = torch.zeros(1, 4, dtype=torch.long)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(ids, ids, torch.zeros(1, 4, 768), torch.ones(1, 4, dtype=torch.bool), torch.zeros(1, 2, 1536),
          torch.ones(1, 2, dtype=torch.bool))
