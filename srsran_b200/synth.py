"""Synthetic input generation for benchmarks and examples: LTE turbo encoding, CRC attachment, code block
segmentation and rate matching in numpy (transmit side, TS 36.212 5.1.1-5.1.4), vectorised over many blocks.

Not part of the decode path and independent of oracle/: it only needs the host-side tables of the library
(QPP permutation, rate-matching order).  tests/test_synth.py checks it against the oracle's encoder chain.
"""
import numpy as np

import srsran_b200 as b

CRC24A = 0x1864CFB
CRC24B = 0x1800063


def _crc24_table(poly):
    t = np.zeros(256, np.uint32)
    for i in range(256):
        c = i << 16
        for _ in range(8):
            c = ((c << 1) ^ poly) if (c & 0x800000) else (c << 1)
        t[i] = c & 0xFFFFFF
    return t


_TAB = {}


def crc24(data, poly):
    """data: (M, nbytes) uint8 -> (M,) CRC24 values, MSB-first, zero init (vectorised over M)"""
    if poly not in _TAB:
        _TAB[poly] = _crc24_table(poly)
    tab = _TAB[poly]
    data = np.atleast_2d(data)
    crc = np.zeros(data.shape[0], np.uint32)
    for i in range(data.shape[1]):
        crc = ((crc << 8) ^ tab[((crc >> 16) & 0xFF) ^ data[:, i]]) & 0xFFFFFF
    return crc


def _rsc(bits):
    """bits: (M, K) -> parity (M, K), final state (M, 3): g0 = 1+D^2+D^3 (feedback), g1 = 1+D+D^3"""
    M, K = bits.shape
    s0 = np.zeros(M, np.uint8)
    s1 = np.zeros(M, np.uint8)
    s2 = np.zeros(M, np.uint8)
    par = np.zeros((M, K), np.uint8)
    for k in range(K):
        fb = bits[:, k] ^ s1 ^ s2
        par[:, k] = fb ^ s0 ^ s2
        s2, s1, s0 = s1, s0, fb
    return par, (s0, s1, s2)


def _tail(state):
    s0, s1, s2 = state
    out = []
    for _ in range(3):
        bit = s1 ^ s2
        fb = bit ^ s1 ^ s2
        par = fb ^ s0 ^ s2
        out += [bit, par]
        s2, s1, s0 = s1, s0, fb
    return np.stack(out, axis=1)  # (M, 6): x, z, x, z, x, z


def turbo_encode(bits):
    """bits: (M, K) uint8 0/1 -> (M, 3K+12) code words in the reference's order: (x_k, z_k, z'_k) triples + tails"""
    bits = np.atleast_2d(bits).astype(np.uint8)
    M, K = bits.shape
    fwd, _ = b.qpp_table(K, 1)
    p1, st1 = _rsc(bits)
    p2, st2 = _rsc(bits[:, fwd])
    cw = np.zeros((M, 3 * K + 12), np.uint8)
    cw[:, 0:3 * K:3] = bits
    cw[:, 1:3 * K:3] = p1
    cw[:, 2:3 * K:3] = p2
    cw[:, 3 * K:3 * K + 6] = _tail(st1)
    cw[:, 3 * K + 6:] = _tail(st2)
    return cw


def encode_tbs(data, tbs, Qm, G, rv=0):
    """data: (M, tbs/8) uint8 -> e-bits (M, G) uint8 0/1 (TB CRC24A, segmentation, CB CRC24B, turbo code, rate matching)"""
    data = np.atleast_2d(data)
    M = data.shape[0]
    rc, seg = b.cbsegm(tbs)
    assert rc == 0 and seg["F"] == 0, "TBS needs filler bits"
    C_, K1, K2, C2 = seg["C"], seg["K1"], seg["K2"], seg["C2"]
    a = crc24(data, CRC24A)
    tb = np.concatenate([data, np.stack([(a >> 16) & 255, (a >> 8) & 255, a & 255], axis=1).astype(np.uint8)], axis=1)
    Gp = G // Qm
    gamma = Gp % C_
    e = np.zeros((M, G), np.uint8)
    rp = wp = 0
    for i in range(C_):
        K = K2 if i < C2 else K1
        rlen = K - 24 if C_ > 1 else K
        n_e = Qm * (Gp // C_) if i <= C_ - gamma - 1 else Qm * ((Gp + C_ - 1) // C_)
        cb = tb[:, rp // 8:(rp + rlen) // 8]
        if C_ > 1:
            c = crc24(cb, CRC24B)
            cb = np.concatenate([cb, np.stack([(c >> 16) & 255, (c >> 8) & 255, c & 255], axis=1).astype(np.uint8)], axis=1)
        cw = turbo_encode(np.unpackbits(cb, axis=1))
        tab = b.rm_table(K, rv, 0)
        idx = tab[np.arange(n_e) % (3 * K + 12)]
        e[:, wp:wp + n_e] = cw[:, idx]
        rp += rlen
        wp += n_e
    return e


def awgn_llr(rng, bits, amp, sigma, dtype):
    """(bit ? +1 : -1) + sigma*n scaled by amp and rounded: the LLR convention of turbodecoder_test.c:246-253"""
    x = amp * ((2.0 * bits.astype(np.float32) - 1.0) + sigma * rng.standard_normal(bits.shape, dtype=np.float32))
    info = np.iinfo(dtype)
    return np.clip(np.rint(x), info.min + 1, info.max).astype(dtype)


# ---------------------------------------------------------------------------------------------- modulation mapper
_AMP = {2: np.array([1, 3]), 3: np.array([3, 1, 5, 7]), 4: np.array([5, 7, 3, 1, 11, 9, 13, 15])}


def lte_modulate(bits, mod):
    """TS 36.211 7.1 modulation mapper: bits (one per element, length = nsym * Qm) -> complex64 symbols.
    mod: 0 BPSK, 1 QPSK, 2 16QAM, 3 64QAM, 4 256QAM (srslte_mod_t)."""
    bits = np.asarray(bits, np.int64)
    if mod == 0:
        return ((1 - 2 * bits) * (1 + 1j) / np.sqrt(2)).astype(np.complex64)
    Qm = 2 * mod
    b_ = bits.reshape(-1, Qm)
    if mod == 1:
        return (((1 - 2 * b_[:, 0]) + 1j * (1 - 2 * b_[:, 1])) / np.sqrt(2)).astype(np.complex64)
    h = Qm // 2 - 1  # amplitude bits per axis
    norm = {2: np.sqrt(10), 3: np.sqrt(42), 4: np.sqrt(170)}[mod]
    ii = np.zeros(len(b_), np.int64)
    qq = np.zeros(len(b_), np.int64)
    for k in range(h):
        ii = 2 * ii + b_[:, 2 + 2 * k]
        qq = 2 * qq + b_[:, 3 + 2 * k]
    amp = _AMP[mod]
    re = (1 - 2 * b_[:, 0]) * amp[ii]
    im = (1 - 2 * b_[:, 1]) * amp[qq]
    return ((re + 1j * im) / norm).astype(np.complex64)


# ---------------------------------------------------------------------------------------------- PUSCH multiplexing (inputs only)
def ul_uci_positions(is_ri, n, Qm, H_prime_total, n_pusch_symbs):
    """first q_bits position of coded ACK / RI symbols 0..n-1 (TS 36.212 5.2.2.8; reference: uci.c:551-605)"""
    rows = H_prime_total // n_pusch_symbs
    norm = n_pusch_symbs > 10
    cols = np.array(([1, 4, 7, 10] if norm else [0, 3, 5, 8]) if is_ri else ([2, 3, 8, 9] if norm else [1, 2, 6, 7]))
    r = np.arange(n)
    return (rows - 1 - r // 4) * Qm + rows * cols[(3 * r) % 4] * Qm


def ul_interleave(rng, g_bits, Qm, rows, n_pusch_symbs, q_ack, q_ri):
    """UL-SCH channel interleaver, transmit side: g_bits (CQI + UL-SCH bits, (M, n)) row by row over the positions without RI;
    RI and then ACK positions carry random bits.  Returns (M, rows * n_pusch_symbs * Qm) bits in transmission order."""
    g_bits = np.atleast_2d(g_bits)
    H = rows * n_pusch_symbs
    k = np.arange(Qm)
    pos = (np.arange(rows)[:, None, None] * Qm + np.arange(n_pusch_symbs)[None, :, None] * rows * Qm + k[None, None, :]).reshape(-1)
    ri_pos = (ul_uci_positions(True, q_ri, Qm, H, n_pusch_symbs)[:, None] + k).reshape(-1)
    ack_pos = (ul_uci_positions(False, q_ack, Qm, H, n_pusch_symbs)[:, None] + k).reshape(-1)
    q = np.zeros((g_bits.shape[0], H * Qm), np.uint8)
    q[:, pos[~np.isin(pos, ri_pos)]] = g_bits
    q[:, ri_pos] = rng.integers(0, 2, (g_bits.shape[0], len(ri_pos)), dtype=np.uint8)
    q[:, ack_pos] = rng.integers(0, 2, (g_bits.shape[0], len(ack_pos)), dtype=np.uint8)
    return q
