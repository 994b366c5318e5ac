"""CPU: the counter-based RNG the dropout contract is written against (oracle/philox.py) -- known-answer tests of
Philox4x32-10 from the Random123 distribution (kat_vectors: zero / all-ones / pi-digits inputs) and the mask statistics."""
import numpy as np

import philox


def _one(ctr, key):
    return [int(x) for x in philox.philox4x32(np.array([ctr], dtype=np.uint32), key)[0]]


def test_known_answer_vectors():
    assert _one([0, 0, 0, 0], (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _one([0xffffffff] * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _one([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], (0xa4093822, 0x299f31d0)) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_mask_contract():
    shape = (37, 101, 64)
    a = philox.keep_mask(shape, 0.1, seed=1234567890123, site=5, step=17)
    b = philox.keep_mask(shape, 0.1, seed=1234567890123, site=5, step=17)
    assert np.array_equal(a, b)                                           # counter-based: reproducible anywhere
    assert abs(a.mean() - 0.9) < 3e-3                                     # 239k draws: sigma = 6e-4
    assert not np.array_equal(a, philox.keep_mask(shape, 0.1, 1234567890123, 6, 17))   # other site
    assert not np.array_equal(a, philox.keep_mask(shape, 0.1, 1234567890123, 5, 18))   # other step
    # a prefix of a longer tensor has the same mask (element index, not shape, drives the stream)
    c = philox.keep_mask((37 * 101 * 64 + 5,), 0.1, 1234567890123, 5, 17)
    assert np.array_equal(c[:a.size], a.reshape(-1))
    x = np.ones(shape, np.float32)
    y, keep = philox.dropout(x, 0.1, 1234567890123, 5, 17)
    assert np.array_equal(keep, a) and np.allclose(y[keep], 1.0 / 0.9) and np.all(y[~keep] == 0)
    assert abs(y.mean() - 1.0) < 4e-3                                     # unbiased


def _host_binary(tmp_path_factory=None):
    """compile tests/philox_host_check.cu for the HOST with nvcc (no GPU needed): it executes csrc/philox.cuh, the source the
    kernels include"""
    import os
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        return None
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "..", "oracle", "_build", "philox_host_check")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(here, "philox_host_check.cu")
    hdr = os.path.join(here, "..", "vl-bert_b200", "csrc", "philox.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call([nvcc, "-O1", "-o", out, src])
    return out


def test_device_source_matches_the_oracle_when_run_on_the_host():
    import subprocess
    import pytest
    exe = _host_binary()
    if exe is None:
        pytest.skip("nvcc not available")
    got = subprocess.check_output([exe, "0x243f6a88", "0x85a308d3", "0x13198a2e", "0x03707344", "0xa4093822", "0x299f31d0"]).decode().split()
    assert got == ["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]
    for (n, p, seed, site, step) in ((1003, 0.1, 1234567890123, 5, 17), (64, 0.5, 7, 0, 0), (257, 0.0, 99, 37, 123456)):
        s = subprocess.check_output([exe, str(n), repr(p), str(seed), str(site), str(step)]).decode().strip()
        ref = philox.keep_mask((n,), p, seed, site, step)
        assert s == "".join("1" if k else "0" for k in ref)
