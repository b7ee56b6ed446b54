"""Shared helpers for the test-suite: seeded synthetic inputs shaped like the reference's tests."""
import numpy as np

SB_PAD = 32


def lanes16(K):
    return 16 if (K % 16 == 0 and K > 800) else (8 if (K % 8 == 0 and K > 400) else 0)


def lanes8(K):
    return 32 if (K % 32 == 0 and K > 2048) else lanes16(K)


def all_K():
    return list(range(40, 513, 8)) + list(range(528, 1025, 16)) + list(range(1056, 2049, 32)) + list(range(2112, 6145, 64))


def random_llr(rng, n, amp, dtype):
    info = np.iinfo(dtype)
    return rng.integers(max(-amp, info.min), min(amp, info.max) + 1, n).astype(dtype)


def bpsk_awgn_llr(rng, bits, amp, sigma, dtype):
    """(bit ? +1 : -1) + sigma n, scaled like turbodecoder_test.c:246-253 (llr_s = 100 * llr)"""
    x = amp * ((2.0 * bits.astype(np.float64) - 1.0) + sigma * rng.standard_normal(len(bits)))
    info = np.iinfo(dtype)
    return np.clip(np.round(x), info.min + 1, info.max).astype(dtype)


def std_to_sb(llr_std, K, N):
    """standard 3k+s order -> the lane layout written by srslte_rm_turbo_rx_lut (rm_turbo.c:263-277)"""
    out = np.zeros(3 * (K + SB_PAD) + 12, llr_std.dtype)
    n = np.arange(K)
    j = (n % (K // N)) * N + n // (K // N)
    for s in range(3):
        out[s * (K + SB_PAD) + j] = llr_std[3 * n + s]
    out[3 * (K + SB_PAD):] = llr_std[3 * K:]
    return out
