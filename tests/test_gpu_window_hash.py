"""The key of the f1 window cache (hash_genome_window, common/util.h:220-241; f1-wrapper.h:97-134 takes it modulo
2^20): the device reads the window a word at a time, the oracle code by code like the reference.  A wrong key would
rarely change a mapping (only which windows share a cache slot), so it gets its own bit-exact test."""
import ctypes as C

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def test_device_window_hash_equals_oracle(gpu_ctx):
    from shrimp_b200._lib import check, lib
    rng = np.random.default_rng(11)
    n_words = 4096
    codes = rng.integers(0, 4, size=n_words * 8, dtype=np.uint32)
    codes[rng.random(codes.size) < 0.02] = 15                       # N: only the low two bits enter the key
    genome = np.zeros(n_words, dtype=np.uint32)
    for j in range(8):
        genome |= codes[j::8] << np.uint32(4 * j)
    lens = np.concatenate([np.arange(0, 70), rng.integers(1, 400, size=3000)]).astype(np.int32)
    offs = rng.integers(0, n_words * 8 - 400, size=lens.size).astype(np.uint32)
    offs[:70] = np.arange(70) % 8 + 8 * rng.integers(0, 100, size=70)   # every length at every misalignment class
    offs[-1], lens[-1] = n_words * 8 - 33, 33                           # a window that ends with the genome
    offs[-2], lens[-2] = n_words * 8 - 16, 16
    got = np.zeros(lens.size, dtype=np.uint32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib().shrimp_gpu_hash_windows(gpu_ctx._h, vp(genome), n_words, vp(offs), vp(lens), lens.size, vp(got)),
          "shrimp_gpu_hash_windows")
    for t in range(lens.size):
        assert int(got[t]) == oracle.hash_genome_window(genome, int(offs[t]), int(lens[t])), (t, offs[t], lens[t])
