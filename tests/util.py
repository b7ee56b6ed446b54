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


# ---- PUSCH grants used by the UL pre-step tests: (tbs, Qm, L_prb, nof_symb, nof_ack, ri_len, cqi kind, I_ack, I_ri, I_cqi);
# cqi kind 0 none / 1 wideband (4 bits) / 2 higher-layer subband N=9 (22 bits).  They cover QPSK/16QAM/64QAM, normal and
# extended CP with and without SRS (12 / 11 / 10 / 9 symbols), 1, 2 and >2 ACK bits, RI, short and long CQI.
UL_GRANTS = [
    (2216, 2, 10, 12, 0, 0, 0, 9, 6, 6),
    (2216, 2, 10, 12, 1, 0, 0, 9, 6, 6),
    (5160, 2, 25, 12, 1, 1, 1, 9, 6, 6),
    (3624, 2, 25, 10, 3, 1, 1, 10, 8, 7),
    (3624, 2, 25, 9, 2, 1, 2, 9, 6, 6),
    (9912, 4, 25, 12, 2, 1, 1, 11, 12, 15),
    (9912, 4, 25, 11, 2, 1, 2, 0, 0, 2),
    (1000, 4, 3, 11, 1, 1, 0, 9, 6, 6),
    (1000, 2, 6, 12, 4, 1, 1, 14, 12, 12),
    (15264, 6, 25, 12, 2, 1, 2, 9, 6, 6),
    (36696, 6, 50, 12, 2, 1, 2, 9, 6, 6),
    (75376, 6, 100, 12, 1, 1, 1, 9, 6, 6),
]
UL_CQI_LEN = {0: 0, 1: 4, 2: 22}


def ul_params(grant, rv=0):
    tbs, Qm, L_prb, nof_symb, nof_ack, ri_len, cqi, I_ack, I_ri, I_cqi = grant
    return (tbs, Qm, L_prb, nof_symb, rv, nof_ack, ri_len, cqi, I_ack, I_ri, I_cqi)


def ul_qprime(port, grant):
    """(Q'_ack, Q'_ri, Q'_cqi) of a grant through the oracle's restatement of uci.c:329-345, 606-630"""
    tbs, Qm, L_prb, nof_symb, nof_ack, ri_len, cqi, I_ack, I_ri, I_cqi = grant
    _, seg = port.cbsegm(tbs)
    K_segm = seg["C1"] * seg["K1"] + seg["C2"] * seg["K2"]
    return port.ulsch_qprime(K_segm, L_prb, nof_symb, nof_ack, ri_len, UL_CQI_LEN[cqi], I_ack, I_ri, I_cqi)


def ul_interleave_tx(port, rng, g_tx, Qm, rows, nsym, qa, qr):
    """Transmit side of the UL-SCH channel interleaver (TS 36.212 5.2.2.8) for the tests: g_tx (CQI + UL-SCH bits, one per
    element) is written row by row over the positions that hold no RI, RI positions and then ACK positions get random bits.
    Positions come from the oracle's restatement of uci.c:551-605.  Returns the H' * Qm interleaved bits."""
    H = rows * nsym
    pos = (np.arange(rows)[:, None, None] * Qm + np.arange(nsym)[None, :, None] * rows * Qm + np.arange(Qm)[None, None, :]).reshape(-1)
    ri_pos = np.concatenate([port.ulsch_uci_position(True, i, Qm, H, nsym) + np.arange(Qm) for i in range(qr)]) if qr else np.zeros(0, np.int64)
    ack_pos = np.concatenate([port.ulsch_uci_position(False, i, Qm, H, nsym) + np.arange(Qm) for i in range(qa)]) if qa else np.zeros(0, np.int64)
    q_tx = np.zeros(H * Qm, np.uint8)
    q_tx[pos[~np.isin(pos, ri_pos)]] = g_tx
    q_tx[ri_pos] = rng.integers(0, 2, len(ri_pos), dtype=np.uint8)
    q_tx[ack_pos] = rng.integers(0, 2, len(ack_pos), dtype=np.uint8)
    return q_tx
