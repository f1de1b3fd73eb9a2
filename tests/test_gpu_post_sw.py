"""post_sw on the device needs the reference's libm bits (sw-post.c ties are broken by the last bit of a sum of
exp() terms): the device's exp / log transcription (shrimp_b200/csrc/glibc_math.cuh) against the host's libm."""
import ctypes as C
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_exp_log_equal_host_libm_bit_for_bit(gpu_ctx):
    from shrimp_b200._lib import check, lib
    rng = np.random.default_rng(5)
    n = 400_000
    x = np.concatenate([
        -rng.random(n) * 60.0,                      # the recurrences' arguments
        -rng.random(n) * 800.0,                     # down to the underflow range
        (rng.random(n) - 0.5) * 1500.0,
        rng.random(n) * 16.0,                       # log arguments: sums of up to 16 terms <= 1
        0.9 + rng.random(n) * 0.2,                  # log near 1 (its second code path)
        np.exp(-rng.random(n) * 700.0),
        rng.integers(0, 2**63 - 1, size=n, dtype=np.int64).view(np.float64),   # any positive bit pattern
        np.array([0.0, -0.0, 1.0, np.inf, -np.inf, 1e-310, 4.9e-324, -1.0, 2.0 ** -1022, 709.78, 709.79, -745.13,
                  -745.14, -708.4, -1022.0, 512.0, -512.0, 10.0, 0.25, 0.75 / 3.0]),
    ]).astype(np.float64)
    e = np.empty_like(x)
    lg = np.empty_like(x)
    check(lib().shrimp_gpu_glibc_explog(gpu_ctx._h, x.ctypes.data_as(C.c_void_p), x.size, e.ctypes.data_as(C.c_void_p),
                                        lg.ctypes.data_as(C.c_void_p)), "shrimp_gpu_glibc_explog")
    libm = C.CDLL("libm.so.6")
    libm.exp.restype = libm.log.restype = C.c_double
    libm.exp.argtypes = libm.log.argtypes = [C.c_double]
    sample = np.concatenate([np.arange(0, x.size, 17), np.arange(x.size - 20, x.size)])
    bad = 0
    for i in sample:
        he, hl = libm.exp(float(x[i])), libm.log(float(x[i]))
        for got, want in ((e[i], he), (lg[i], hl)):
            if not (np.float64(got).view(np.uint64) == np.float64(want).view(np.uint64) or (got != got and want != want)):
                bad += 1
                if bad < 5:
                    print("DIFF", float(x[i]).hex(), float(got).hex(), float(want).hex())
    assert bad == 0
