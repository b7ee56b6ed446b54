"""GPU parity tests: the CUDA path (through the C ABI of libsrslte_fec_b200.so) against the CPU oracle on the
same seeded inputs.  Bit-exact: decided bytes after every half-iteration, CRC outcomes, iteration counts,
soft-buffer contents.  Marked gpu: run on the B200 box."""
import ctypes as C
import os

import numpy as np
import pytest

import srsran_b200 as b
from oracle.bindings import (CRC8, CRC16, CRC24A, CRC24B, TDEC_AUTO, TDEC_AVX8_WINDOW, TDEC_AVX_WINDOW, TDEC_GENERIC, TDEC_SSE8_WINDOW,
                             TDEC_SSE_WINDOW)
from util import all_K, bpsk_awgn_llr, lanes8, lanes16, random_llr, std_to_sb

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------------------- CRC (D1)
def test_crc_drop_in(port):
    rng = np.random.default_rng(11)
    for n in (1, 2, 3, 5, 31, 32, 33, 64, 100, 728, 768, 1000, 9422, 12237):
        d = rng.integers(0, 256, n, dtype=np.uint8)
        for poly, order in ((CRC24A, 24), (CRC24B, 24), (CRC16, 16), (CRC8, 8)):
            assert b.crc_checksum_byte(poly, order, d) == port.crc_bytes(poly, order, d), (n, hex(poly))
    for n in (5001, 40, 41, 47, 6144):
        bits = rng.integers(0, 2, n, dtype=np.uint8)
        for poly, order in ((CRC24A, 24), (CRC24B, 24), (CRC16, 16), (CRC8, 8)):
            assert b.crc_checksum_bits(poly, order, bits) == port.crc_bits(poly, order, bits), (n, hex(poly))


def test_crc_zero_and_codeword(port):
    # a message followed by its own CRC checks to zero (the early-stop criterion, sch.c:441)
    rng = np.random.default_rng(12)
    d = rng.integers(0, 256, 765, dtype=np.uint8)
    c = port.crc_bytes(CRC24B, 24, d)
    full = np.concatenate([d, np.array([(c >> 16) & 255, (c >> 8) & 255, c & 255], np.uint8)])
    assert b.crc_checksum_byte(CRC24B, 24, full) == 0
    assert b.crc_checksum_byte(CRC24A, 24, np.zeros(100, np.uint8)) == 0


# ----------------------------------------------------------------------------------------- rate de-matching (B1)
@pytest.mark.parametrize("K,E", [(40, 132), (40, 500), (504, 1000), (1024, 1000), (5824, 6918), (5824, 6924), (6144, 7200), (6144, 18444), (6144, 30000),
                                 (2048, 6156), (2112, 3000), (816, 2460)])
def test_rm_rx_lut(port, K, E):
    rng = np.random.default_rng(K + E)
    ci = port.cbindex(K)
    for rv in range(4):
        e = random_llr(rng, E, 32767, np.int16)
        for sb in (True, False):
            a = random_llr(rng, 18600, 32767, np.int16)
            want = a.copy()
            port.rm_rx16(e, want, K, rv, lanes16(K) if sb else 0)
            assert b.rm_turbo_rx_lut(e, a, ci, rv, sb) == 0
            assert (a == want).all(), (K, E, rv, sb)
        e8 = random_llr(rng, E, 127, np.int8)
        a = random_llr(rng, 18600 * 2, 127, np.int8)
        want = a.copy()
        port.rm_rx8(e8, want, K, rv, lanes8(K))
        assert b.rm_turbo_rx_lut(e8, a, ci, rv) == 0
        assert (a == want).all(), (K, E, rv, "int8")


def test_rm_rx_lut_invalid():
    e = np.zeros(100, np.int16)
    o = np.zeros(18600, np.int16)
    assert b.rm_turbo_rx_lut(e, o, 188, 0) == -2
    assert b.rm_turbo_rx_lut(e, o, 0, 4) == -2


# ----------------------------------------------------------------------------------------- decoder, per half-iteration (C1-C6)
def _tdec_case(port, K, dtype, amp, nit, dec_type=TDEC_AUTO, force_not_sb=False, seed=0):
    rng = np.random.default_rng(seed + K)
    bits = 16 if dtype == np.int16 else 8
    if dec_type == TDEC_AUTO:
        N = lanes16(K) if bits == 16 else lanes8(K)
        sb_in = N > 0 and not force_not_sb
    else:
        sb_in = False
    n = 3 * (K + 32) + 12 if sb_in else 3 * K + 12
    llr = random_llr(rng, n, amp, dtype)
    hp = port.tdec_new(dec_type, force_not_sb)
    assert port.tdec_new_cb(hp, K) == 0
    td = b.TurboDecoder(6144, dec_type, force_not_sb)
    assert td.new_cb(K) == 0
    for it in range(nit):
        want = port.tdec_iteration(hp, llr, K)
        got = td.iteration(llr, K)
        assert td.n_iter() == it + 1
        assert (got == want).all(), "K=%d bits=%d amp=%d dec=%d half-iteration %d: %d byte mismatches" % (K, bits, amp, dec_type, it, int((got != want).sum()))
    td.free()
    port.tdec_del(hp)


@pytest.mark.parametrize("K,amp,nit", [(6144, 300, 4), (6144, 30000, 6), (816, 5000, 6), (512, 20000, 6), (408, 300, 5), (800, 32767, 4), (5824, 2000, 10),
                                       (1008, 700, 4), (2048, 100, 4), (2112, 32767, 3), (4160, 900, 5)])
def test_tdec_int16_windowed(port, K, amp, nit):
    _tdec_case(port, K, np.int16, amp, nit)


@pytest.mark.parametrize("K,amp,nit", [(6144, 60, 6), (6144, 127, 6), (2112, 100, 5), (816, 127, 6), (2048, 90, 6), (1008, 127, 4), (5824, 50, 10), (3072, 30, 4)])
def test_tdec_int8_windowed(port, K, amp, nit):
    _tdec_case(port, K, np.int8, amp, nit)


@pytest.mark.parametrize("K,amp,nit", [(40, 100, 6), (400, 300, 6), (400, 20000, 6), (200, 32767, 5), (104, 1000, 10), (48, 50, 3)])
def test_tdec_generic(port, K, amp, nit):
    _tdec_case(port, K, np.int16, amp, nit)


@pytest.mark.parametrize("K,amp,nit", [(40, 100, 4), (400, 127, 6), (408, 127, 4), (512, 60, 6), (800, 127, 5)])
def test_tdec_int8_small_K(port, K, amp, nit):
    # int8 input, K <= 800: widened to int16 then generic / 8-lane rules (408..800 is the documented deviation:
    # full conversion instead of the reference's uninitialised read)
    _tdec_case(port, K, np.int8, amp, nit)


@pytest.mark.parametrize("K,dec,fnsb", [(6144, TDEC_AVX_WINDOW, True), (504, TDEC_SSE_WINDOW, True), (504, TDEC_GENERIC, True), (6144, TDEC_GENERIC, False),
                                        (1024, TDEC_SSE_WINDOW, False), (6144, TDEC_SSE_WINDOW, True), (6144, TDEC_AVX8_WINDOW, True),
                                        (1024, TDEC_SSE8_WINDOW, True), (6144, TDEC_SSE8_WINDOW, True)])
def test_tdec_manual(port, K, dec, fnsb):
    _tdec_case(port, K, np.int16, 400, 4, dec, fnsb)


@pytest.mark.parametrize("K", [504, 6144, 40, 2048])
def test_tdec_auto_force_not_sb(port, K):
    _tdec_case(port, K, np.int16, 400, 4, TDEC_AUTO, True)


def test_tdec_errors():
    td = b.TurboDecoder(6144)
    assert td.new_cb(6145) == -1
    assert td.new_cb(41) == -1  # not an LTE size (stricter than the reference)
    td.free()
    td = b.TurboDecoder(1024)
    assert td.new_cb(2048) == -1
    # iteration without new_cb is a silent no-op (turbodecoder.c:530)
    out = td.iteration(np.zeros(3 * 1024 + 12 + 96, np.int16), 1024)
    assert td.n_iter() == 0 and not out.any()
    td.free()
    with pytest.raises(b.B200Error):
        b.TurboDecoder(6144, 2)  # SRSLTE_TDEC_SSE (state-parallel) is not provided


# ----------------------------------------------------------------------------------------- all 188 sizes (config C3)
def test_tdec_all_sizes_batch(port, ctx):
    rng = np.random.default_rng(188)
    for dtype, amp in ((np.int16, 600), (np.int8, 40)):
        for K in all_K():
            bits = rng.integers(0, 2, K, dtype=np.uint8)
            cw = port.tcod_encode(bits)
            llr = bpsk_awgn_llr(rng, cw, amp / 4, 0.8, dtype)
            N = lanes16(K) if dtype == np.int16 else lanes8(K)
            x = std_to_sb(llr, K, N) if N else llr
            batch = np.ascontiguousarray(np.stack([x, x]))
            got = ctx.tdec_batch(batch, K, 3, input_sb=N > 0)
            hp = port.tdec_new(TDEC_AUTO, False)
            rc, want = port.tdec_run_all(hp, x, 3, K)
            port.tdec_del(hp)
            assert rc == 0
            assert (got[0] == want).all() and (got[1] == want).all(), (K, dtype)


# ----------------------------------------------------------------------------------------- batched code blocks (config C1)
def test_cb_batch_c1(port, ctx):
    rng = np.random.default_rng(61)
    K, ncb = 6144, 150
    llr = np.zeros((ncb, 3 * K + 12), np.int16)
    data = []
    for i in range(ncb):
        bits = rng.integers(0, 2, K, dtype=np.uint8)
        data.append(bits)
        llr[i] = bpsk_awgn_llr(rng, port.tcod_encode(bits), 100, 0.9 + 0.4 * (i % 5) / 4, np.int16)
    got = ctx.tdec_batch(llr, K, 4, input_sb=False)
    hp = port.tdec_new(TDEC_AUTO, True)
    n_ok = 0
    for i in range(ncb):
        rc, want = port.tdec_run_all(hp, llr[i], 4, K)
        assert (got[i] == want).all(), i
        n_ok += int((np.unpackbits(want) == data[i]).all())
    port.tdec_del(hp)
    assert n_ok > ncb // 4  # the synthetic SNR range really decodes a good share of blocks
    assert ctx.last_launches() > 0 and ctx.last_map_launches() == 4


def test_cb_batch_saturating_and_int8(port, ctx):
    rng = np.random.default_rng(62)
    for K, dtype, amp, nit in ((6144, np.int16, 32767, 6), (6144, np.int8, 127, 5), (1024, np.int16, 20000, 4), (504, np.int16, 32767, 4), (120, np.int16, 32767, 5)):
        N = lanes16(K) if dtype == np.int16 else lanes8(K)
        n = 3 * (K + 32) + 12 if N else 3 * K + 12
        llr = np.stack([random_llr(rng, n, amp, dtype) for _ in range(21)])
        got = ctx.tdec_batch(llr, K, nit, input_sb=N > 0)
        hp = port.tdec_new(TDEC_AUTO, False)
        for i in range(llr.shape[0]):
            rc, want = port.tdec_run_all(hp, llr[i], nit, K)
            assert (got[i] == want).all(), (K, dtype, i)
        port.tdec_del(hp)


# ----------------------------------------------------------------------------------------- transport blocks (A1, A2, E2; configs C2, C4)
def _tb_inputs(port, rng, tbs, Qm, G, rv, dtype, amp, sigma, data=None):
    if data is None:
        data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
    e = port.encode_tb(tbs, Qm, rv, G, data)
    return data, bpsk_awgn_llr(rng, e, amp, sigma, dtype)


def _decode_both(port, ctx, tbs, Qm, rv, llr, max_iter, sb_port, sb_gpu):
    rc_p, d_p, nit_p, avg_p, crc_p = port.decode_tb(sb_port, tbs, Qm, rv, llr, max_iter)
    t = b.make_tbs(1)
    out = np.zeros(tbs // 8 + 6 + 16, np.uint8)
    t[0].e_bits = llr.ctypes.data
    t[0].nof_e_bits = len(llr)
    t[0].tbs = tbs
    t[0].Qm = Qm
    t[0].rv = rv
    t[0].softbuffer = sb_gpu
    t[0].data = out.ctypes.data
    ctx.decode_tbs(t, llr.dtype == np.int8, max_iter)
    C_ = t[0].nof_cb
    n = tbs // 8 + (3 if C_ == 1 else 6)
    assert t[0].ret == rc_p
    assert (out[:n] == d_p[:n]).all()
    assert list(t[0].cb_crc[:C_]) == crc_p[:C_].tolist()
    assert list(t[0].cb_noi[:C_]) == nit_p[:C_].tolist()
    assert abs(t[0].avg_iterations - avg_p) < 1e-6
    return rc_p, avg_p


@pytest.mark.parametrize("tbs,Qm,G,dtype,amp,sigma,max_iter", [
    (75376, 6, 90000, np.int16, 100, 0.0, 8), (75376, 6, 90000, np.int16, 100, 0.46, 8), (75376, 6, 90000, np.int16, 100, 0.52, 8), (75376, 6, 90000, np.int16, 100, 3.0, 4),
    (97896, 8, 115200, np.int8, 20, 0.42, 8), (6120, 2, 14400, np.int16, 50, 0.9, 8), (1000, 2, 2880, np.int16, 50, 0.9, 8),
    (296, 2, 1200, np.int8, 30, 0.5, 8), (15264, 4, 20000, np.int16, 200, 0.55, 10), (40, 2, 480, np.int16, 80, 0.5, 4),
    (2216, 2, 7000, np.int8, 30, 0.6, 6)])
def test_decode_tb(port, ctx, tbs, Qm, G, dtype, amp, sigma, max_iter):
    rng = np.random.default_rng(tbs + G)
    sbp = port.softbuffer_new()
    data, llr = _tb_inputs(port, rng, tbs, Qm, G, 0, dtype, amp, sigma)
    _decode_both(port, ctx, tbs, Qm, 0, llr, max_iter, sbp, None)
    port.softbuffer_del(sbp)


def test_decode_tb_harq(port, ctx):
    """HARQ: failed first transmission, soft-combining of retransmissions, CRC-ok code blocks skipped (sch.c:385,462-484)"""
    rng = np.random.default_rng(77)
    for tbs, Qm, G, dtype, amp, sigma in ((75376, 6, 90000, np.int16, 100, 0.95), (97896, 8, 115200, np.int8, 20, 0.8), (15264, 4, 20000, np.int16, 200, 0.75)):
        sbp = port.softbuffer_new()
        sbg = ctx.softbuffer_create()
        data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        rcs = []
        for rv in (0, 2, 3, 1):
            _, llr = _tb_inputs(port, rng, tbs, Qm, G, rv, dtype, amp, sigma, data)
            rc, avg = _decode_both(port, ctx, tbs, Qm, rv, llr, 8, sbp, sbg)
            rcs.append(rc)
            if rc == 0:
                break
        assert rcs[0] == -1 and rcs[-1] == 0, rcs  # the case really exercises combining
        # reset + new data on the same soft buffer
        port.softbuffer_reset(sbp)
        ctx.softbuffer_reset(sbg)
        _, llr = _tb_inputs(port, rng, tbs, Qm, G, 0, dtype, amp, 0.3)
        rc, _ = _decode_both(port, ctx, tbs, Qm, 0, llr, 8, sbp, sbg)
        assert rc == 0
        port.softbuffer_del(sbp)
        ctx.softbuffer_free(sbg)


def test_decode_tbs_batch_mixed(port, ctx):
    """many TBs of different sizes in one launch, mixed outcomes and iteration counts"""
    rng = np.random.default_rng(99)
    cases = [(75376, 6, 90000, 0.42), (75376, 6, 90000, 0.5), (6120, 2, 14400, 0.9), (1000, 2, 2880, 0.9), (15264, 4, 20000, 0.5), (40, 2, 480, 0.5),
             (75376, 6, 90000, 2.0), (2216, 2, 7000, 0.8), (31704, 6, 40000, 0.5), (4584, 4, 9000, 0.6)] * 2
    n = len(cases)
    t = b.make_tbs(n)
    llrs, outs = [], []
    for i, (tbs, Qm, G, sigma) in enumerate(cases):
        _, llr = _tb_inputs(port, rng, tbs, Qm, G, 0, np.int16, 100, sigma)
        out = np.zeros(tbs // 8 + 22, np.uint8)
        llrs.append(llr)
        outs.append(out)
        t[i].e_bits, t[i].nof_e_bits, t[i].tbs, t[i].Qm, t[i].rv, t[i].data = llr.ctypes.data, len(llr), tbs, Qm, 0, out.ctypes.data
    ctx.decode_tbs(t, False, 8)
    rets = []
    for i, (tbs, Qm, G, sigma) in enumerate(cases):
        sbp = port.softbuffer_new()
        rc, d, nit, avg, crc = port.decode_tb(sbp, tbs, Qm, 0, llrs[i], 8)
        port.softbuffer_del(sbp)
        C_ = t[i].nof_cb
        nb = tbs // 8 + (3 if C_ == 1 else 6)
        assert t[i].ret == rc and (outs[i][:nb] == d[:nb]).all(), i
        assert list(t[i].cb_noi[:C_]) == nit[:C_].tolist(), i
        assert abs(t[i].avg_iterations - avg) < 1e-6
        rets.append(rc)
    assert 0 in rets and -1 in rets


@pytest.mark.parametrize("dtype,amp", [(np.int16, 100), (np.int8, 25)])
def test_decode_tbs_all_sizes_one_batch(port, ctx, dtype, amp):
    """config C3: ONE submit whose code blocks cover all 188 LTE sizes (single-CB transport blocks, tbs = K - 24), CRC
    early stopping with iteration counts that differ per block (noise level varies), plus a few multi-CB blocks in
    between -- every K is its own K-group / tensor map, every decoder class is present"""
    rng = np.random.default_rng(1880 + (dtype == np.int8))
    cases = []
    for i, K in enumerate(all_K()):
        tbs = K - 24
        G = 2 * ((3 * K + 12) // 2 + (i % 5) * 40)  # around the mother code rate, sometimes with repetition
        cases.append((tbs, 2, G, (0.45, 0.7, 0.95, 1.3)[i % 4]))
        if i % 47 == 0:
            cases.append((15264, 4, 20000, 0.5))
    n = len(cases)
    t = b.make_tbs(n)
    llrs, outs = [], []
    for i, (tbs, Qm, G, sigma) in enumerate(cases):
        _, llr = _tb_inputs(port, rng, tbs, Qm, G, 0, dtype, amp, sigma)
        out = np.zeros(tbs // 8 + 22, np.uint8)
        llrs.append(llr)
        outs.append(out)
        t[i].e_bits, t[i].nof_e_bits, t[i].tbs, t[i].Qm, t[i].rv, t[i].data = llr.ctypes.data, len(llr), tbs, Qm, 0, out.ctypes.data
    ctx.decode_tbs(t, dtype == np.int8, 8)
    its = set()
    for i, (tbs, Qm, G, sigma) in enumerate(cases):
        sbp = port.softbuffer_new()
        rc, d, nit, avg, crc = port.decode_tb(sbp, tbs, Qm, 0, llrs[i], 8)
        port.softbuffer_del(sbp)
        C_ = t[i].nof_cb
        nb = tbs // 8 + (3 if C_ == 1 else 6)
        assert t[i].ret == rc and (outs[i][:nb] == d[:nb]).all(), (i, tbs)
        assert list(t[i].cb_noi[:C_]) == nit[:C_].tolist(), (i, tbs)
        its.update(nit[:C_].tolist())
    assert len(its) >= 4  # iteration counts really vary inside the batch


@pytest.mark.gpu
def test_tb_hints_change_the_grouping_not_the_results(port, ctx):
    """srslte_b200_set_tb_hints / automatic grouping: code blocks of equal size are grouped by the caller's difficulty hint, or
    by the engine's own noise estimate when there is none.  240 transport blocks
    of three sizes (8-, 16-lane and generic decoder classes) and four noise levels, decoded without hints, with the noise level
    as hint, with random hints and with a hint array of the wrong length (ignored): identical return codes, bytes, CRC flags and
    half-iteration counts every time, all equal to the oracle's"""
    rng = np.random.default_rng(777)
    shapes = [(6120, 2, 18700), (1000, 2, 3300), (376, 2, 1300)]
    sigmas = (0.5, 0.8, 1.0, 1.4)
    base = {}
    for tbs, Qm, G in shapes:
        for j, sg in enumerate(sigmas):
            _, llr = _tb_inputs(port, rng, tbs, Qm, G, 0, np.int16, 100, sg)
            sbp = port.softbuffer_new()
            base[(tbs, j)] = (llr, port.decode_tb(sbp, tbs, Qm, 0, llr, 8))
            port.softbuffer_del(sbp)
    cases = [(shapes[i % 3], (i * 7 + i // 3) % 4) for i in range(240)]
    n = len(cases)
    runs = []
    for mode in ("none", "sigma", "random", "wrong-length", "no-auto-group"):
        t = b.make_tbs(n)
        outs = np.zeros((n, 6120 // 8 + 22), np.uint8)
        for i, ((tbs, Qm, G), j) in enumerate(cases):
            llr = base[(tbs, j)][0]
            t[i].e_bits, t[i].nof_e_bits, t[i].tbs, t[i].Qm, t[i].rv, t[i].data = llr.ctypes.data, G, tbs, Qm, 0, outs[i].ctypes.data
        if mode == "sigma":
            ctx.set_tb_hints([sigmas[j] for _, j in cases])
        elif mode == "random":
            ctx.set_tb_hints(rng.standard_normal(n))
        elif mode == "wrong-length":
            ctx.set_tb_hints(np.ones(n - 1))
        # (without hints the engine orders the blocks of a size by a noise estimate of its own, k_cb_stat / k_regroup:
        #  "none" and "wrong-length" run that path, "no-auto-group" the plain submission order)
        ctx.set_option("auto_group", 0 if mode == "no-auto-group" else 1)
        ctx.decode_tbs(t, False, 8)
        for i, ((tbs, Qm, G), j) in enumerate(cases):
            rc, d, nit, avg, crc = base[(tbs, j)][1]
            nb = tbs // 8 + 3
            assert t[i].ret == rc and (outs[i][:nb] == d[:nb]).all(), (mode, i)
            assert t[i].cb_noi[0] == nit[0] and t[i].cb_crc[0] == crc[0], (mode, i)
        runs.append((outs.copy(), [t[i].ret for i in range(n)]))
    for o, r in runs[1:]:
        assert (o == runs[0][0]).all() and r == runs[0][1]


def test_decode_tb_invalid(ctx):
    t = b.make_tbs(3)
    out = np.zeros(20000, np.uint8)
    llr = np.zeros(90000, np.int16)
    for i, tbs in enumerate((75376, 0, 75000)):  # ok-shaped, empty, filler bits needed (F > 0)
        t[i].e_bits, t[i].nof_e_bits, t[i].tbs, t[i].Qm, t[i].rv, t[i].data = llr.ctypes.data, 90000, tbs, 6, 0, out.ctypes.data
    ctx.decode_tbs(t, False, 2)
    assert t[0].ret == -1       # all-zero LLRs: CRC fails
    assert t[1].ret == 0        # tbs == 0 -> SRSLTE_SUCCESS (sch.c:515)
    assert t[2].ret == -2       # filler bits not supported (sch.c:519-522)


# ----------------------------------------------------------------------------------------- Fast16 path: monitor + exact replay
def test_fast16_replay_accounting(port, ctx):
    """small LLRs decode entirely on the native packed-instruction path; full-scale LLRs are all replayed with the
    exact saturating arithmetic; either way the bytes equal the oracle's and fast16=0 gives the same bytes"""
    rng = np.random.default_rng(1616)
    K, ncb = 6144, 24
    for amp, nit, expect in ((100, 4, "none"), (32767, 4, "all"), (300, 6, "some")):
        llr = np.zeros((ncb, 3 * K + 12), np.int16)
        for i in range(ncb):
            bits = rng.integers(0, 2, K, dtype=np.uint8)
            llr[i] = bpsk_awgn_llr(rng, port.tcod_encode(bits), amp, 0.9, np.int16) if amp < 32767 else random_llr(rng, 3 * K + 12, amp, np.int16)
        ctx.set_option("fast16", 1)
        got = ctx.tdec_batch(llr, K, nit)
        rep, tot = ctx.last_replayed(), ctx.last_half_iterations()
        assert tot == ncb * nit
        if expect == "none":
            assert rep == 0
        elif expect == "all":
            assert rep == tot
        else:
            assert 0 < rep < tot
        ctx.set_option("fast16", 0)
        exact = ctx.tdec_batch(llr, K, nit)
        assert ctx.last_replayed() == 0
        ctx.set_option("fast16", 1)
        assert (got == exact).all()
        hp = port.tdec_new(TDEC_AUTO, True)
        for i in range(ncb):
            rc, want = port.tdec_run_all(hp, llr[i], nit, K)
            assert (got[i] == want).all(), (amp, i)
        port.tdec_del(hp)


def test_fast16_amplitude_sweep(port, ctx):
    """sweep the LLR amplitude across the range-monitor's decision boundary: bit-exact everywhere, 8 half-iterations"""
    rng = np.random.default_rng(1717)
    for K in (6144, 1024, 512):
        N = lanes16(K)
        for amp in (30, 100, 180, 250, 400, 700, 1500, 4000, 12000):
            bits = rng.integers(0, 2, K, dtype=np.uint8)
            x = std_to_sb(bpsk_awgn_llr(rng, port.tcod_encode(bits), amp, 0.8, np.int16), K, N)
            batch = np.ascontiguousarray(np.stack([x] * 3))
            got = ctx.tdec_batch(batch, K, 8, input_sb=True)
            hp = port.tdec_new(TDEC_AUTO, False)
            rc, want = port.tdec_run_all(hp, x, 8, K)
            port.tdec_del(hp)
            assert (got == want[None, :]).all(), (K, amp)


def test_fast16_all_window_sizes_across_the_monitor_boundary(port, ctx):
    """every windowed int16 size (partial top tiles for W % 8 != 0, 8- and 16-lane decoders), three amplitudes -- far
    below, around and far above the range monitor's decision boundary, so that blocks of one batch take the fast path,
    the exact replay, or both in different half-iterations -- random (non-codeword) LLRs, 5 half-iterations"""
    rng = np.random.default_rng(1919)
    replayed = clean = 0
    for K in all_K():
        N = lanes16(K)
        if N == 0:
            continue
        for amp in (60, 900, 6000):
            batch = np.stack([random_llr(rng, 3 * K + 12, amp, np.int16) for _ in range(3)])
            got = ctx.tdec_batch(batch, K, 5, input_sb=False)
            rep = ctx.last_replayed()
            replayed += rep > 0
            clean += rep == 0
            hp = port.tdec_new(TDEC_AUTO, True)
            for i in range(3):
                rc, want = port.tdec_run_all(hp, batch[i], 5, K)
                assert rc == 0 and (got[i] == want).all(), (K, amp, i, rep)
            port.tdec_del(hp)
    assert replayed > 50 and clean > 50


def test_exact_path_full_suite_sample(port, ctx):
    """the exact-arithmetic kernels alone (fast16 = 0) on the transport-block path"""
    rng = np.random.default_rng(1818)
    ctx.set_option("fast16", 0)
    try:
        for tbs, Qm, G, dtype, amp, sigma in ((75376, 6, 90000, np.int16, 100, 0.46), (15264, 4, 20000, np.int16, 3000, 0.55)):
            sbp = port.softbuffer_new()
            data, llr = _tb_inputs(port, rng, tbs, Qm, G, 0, dtype, amp, sigma)
            _decode_both(port, ctx, tbs, Qm, 0, llr, 8, sbp, None)
            port.softbuffer_del(sbp)
    finally:
        ctx.set_option("fast16", 1)


# ----------------------------------------------------------------------------------------- drop-in decode_tb + srslte_softbuffer_rx_*
def test_drop_in_decode_tb_with_softbuffer(port):
    """srslte_b200_decode_tb (what sch.c's decode_tb becomes) with a srslte_softbuffer_rx_t across HARQ retransmissions"""
    rng = np.random.default_rng(4242)
    for tbs, Qm, G, dtype, amp, sigma in ((15264, 4, 20000, np.int16, 200, 0.75), (31704, 6, 40000, np.int8, 20, 0.62)):
        sbp = port.softbuffer_new()
        sbg = b.SoftbufferRx(100)
        sbg.reset_tbs(tbs)
        data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        rcs = []
        for rv in (0, 2, 3):
            _, llr = _tb_inputs(port, rng, tbs, Qm, G, rv, dtype, amp, sigma, data)
            rc_p, d_p, nit_p, avg_p, crc_p = port.decode_tb(sbp, tbs, Qm, rv, llr, 8)
            rc, d, avg = b.decode_tb(sbg, tbs, Qm, rv, llr, 8)
            C_ = port.cbsegm(tbs)[1]["C"]
            n = tbs // 8 + 3
            assert rc == rc_p and (d[:n] == d_p[:n]).all() and abs(avg - avg_p) < 1e-6
            assert sbg.cb_crc(C_) == [bool(x) for x in crc_p[:C_]]
            rcs.append(rc)
            if rc == 0:
                assert (d[:tbs // 8] == data).all()
                break
        assert rcs[0] == -1 and rcs[-1] == 0, rcs
        sbg.free()
        port.softbuffer_del(sbp)


# ----------------------------------------------------------------------------------------- soft demodulation
@pytest.mark.parametrize("dtype", [np.int16, np.int8])
def test_demod_descramble(port, ctx, dtype):
    """k_demod_descramble against the oracle: every modulation, lengths around the reference's SIMD group sizes (its bodies
    and tails round differently), amplitudes up to saturation, with and without descrambling, many codewords per call"""
    rng = np.random.default_rng(31 + (dtype == np.int8))
    cws, want = [], []
    for mod in range(5):
        for n in (1, 3, 4, 7, 8, 9, 15, 16, 17, 31, 33, 100, 1001, 15000):
            amp = (0.3, 1.0, 3.0, 50.0, 400.0)[(n + mod) % 5]
            sym = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * amp).astype(np.complex64)
            nbits = n * b.MOD_BITS[mod]
            scr = port.sequence_bytes(int(rng.integers(1, 2 ** 31)), nbits) if (n + mod) % 3 else None
            cws.append((sym, mod, scr))
            llr = port.demod(mod, sym, dtype)
            want.append(port.descramble(scr, llr) if scr is not None else llr)
    got = ctx.demod_descramble(cws, dtype)
    for i, (g, w) in enumerate(zip(got, want)):
        assert (g == w).all(), (i, cws[i][1], len(cws[i][0]))


def test_sequence_bytes_helper(port):
    for c_init, L in ((1, 100), (12345, 90000), (0x7fffffff, 115203), (777, 33)):
        assert (b.sequence_bytes(c_init, L) == port.sequence_bytes(c_init, L)).all()


@pytest.mark.parametrize("tbs,mod,dtype,sigma", [(75376, 3, np.int16, 0.08), (97896, 4, np.int8, 0.012), (15264, 2, np.int16, 0.15), (1000, 1, np.int8, 0.3)])
def test_symbols_to_transport_block_on_device(port, ctx, tbs, mod, dtype, sigma):
    """the chain of pdsch.c:832-859 on the device: equalised symbols -> soft demodulation -> descrambling -> decode_tb, LLRs
    never leaving the GPU (demod with OUT_DEVICE, decode with IN_DEVICE), against the same chain through the oracle"""
    from srsran_b200 import synth
    rng = np.random.default_rng(tbs + mod)
    Qm = b.MOD_BITS[mod]
    nsym = {75376: 15000, 97896: 14400, 15264: 5000, 1000: 1440}[tbs]
    G = nsym * Qm
    data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
    e = port.encode_tb(tbs, Qm, 0, G, data)  # rate-matched bits (one per byte)
    c_init = (0x1234 << 14) + (0 << 13) + (3 << 9) + 77
    scr = port.sequence_bytes(c_init, G)
    scr_bits = np.unpackbits(scr)[:G]
    sym = synth.lte_modulate(e ^ scr_bits, mod)
    sym = (sym + sigma * (rng.standard_normal(nsym) + 1j * rng.standard_normal(nsym))).astype(np.complex64)
    # oracle chain
    llr = port.descramble(scr, port.demod(mod, sym, dtype))
    sbp = port.softbuffer_new()
    rc, want, nit, avg, crc = port.decode_tb(sbp, tbs, Qm, 0, llr, 8)
    port.softbuffer_del(sbp)
    # device chain
    d_e = ctx.device_alloc(G * np.dtype(dtype).itemsize + 64)
    dm = b.make_demods(1)
    dm[0].symbols, dm[0].nof_symbols, dm[0].mod, dm[0].scramble_bytes, dm[0].e_bits = sym.ctypes.data, nsym, mod, scr.ctypes.data, d_e
    ctx.demod_descramble_raw(dm, dtype == np.int8, b.OUT_DEVICE)
    d_out = ctx.device_alloc(tbs // 8 + 64)
    t = b.make_tbs(1)
    t[0].e_bits, t[0].nof_e_bits, t[0].tbs, t[0].Qm, t[0].rv, t[0].softbuffer, t[0].data = d_e, G, tbs, Qm, 0, None, d_out
    ctx.decode_tbs(t, dtype == np.int8, 8, flags=b.IN_DEVICE | b.OUT_DEVICE)
    out = np.zeros(tbs // 8 + 6, np.uint8)
    ctx.d2h(out, d_out)
    ctx.device_free(d_e)
    ctx.device_free(d_out)
    C_ = t[0].nof_cb
    nb = tbs // 8 + (3 if C_ == 1 else 6)
    assert t[0].ret == rc and (out[:nb] == want[:nb]).all()
    assert rc == 0 and (out[:tbs // 8] == data).all()  # (the case is meant to decode)
    assert list(t[0].cb_noi[:C_]) == nit[:C_].tolist()


# ----------------------------------------------------------------------------------------- transmit mirror
def test_encode_tbs(port, ctx):
    """srslte_dlsch_encode2 semantics (TB CRC24A, segmentation, CB CRC24B, turbo code, rate matching) for many transport
    blocks in one call, against the oracle's encoder chain (itself checked against the reference's sch.c): single- and
    multi-CB blocks, every rv, puncturing and repetition, e-bit counts that leave code blocks at odd bit offsets"""
    rng = np.random.default_rng(707)
    cases = [(75376, 6, 90000, 0), (75376, 6, 90000, 2), (97896, 8, 115200, 0), (97896, 8, 115200, 3), (15264, 4, 20000, 1), (6120, 2, 14400, 0),
             (1000, 2, 2880, 0), (1000, 2, 9000, 1), (40, 2, 480, 0), (40, 2, 132, 2), (2216, 2, 7000, 3), (31704, 6, 40002, 0), (4584, 4, 9004, 0),
             (296, 2, 1202, 1), (6120, 2, 30000, 2)]
    blocks, want = [], []
    for tbs, Qm, G, rv in cases:
        data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        blocks.append((data, tbs, Qm, rv, G))
        want.append(np.packbits(port.encode_tb(tbs, Qm, rv, G, data)))
    got, rets = ctx.encode_tbs(blocks)
    for i, (g, w) in enumerate(zip(got, want)):
        assert rets[i] == 0
        assert (g == w).all(), (i, cases[i], int(np.argmax(g != w)))
    # a transport block size that needs filler bits is rejected like the reference does (sch.c:249-252)
    _, rets = ctx.encode_tbs([(np.zeros(75000 // 8, np.uint8), 75000, 6, 0, 90000), (np.zeros(496 // 8, np.uint8), 496, 2, 0, 1600)])
    assert rets == [-1, -1]


def test_encode_decode_round_trip_on_device(port, ctx):
    """encode -> (bits as LLRs) -> decode, and the encoder's output decoded by the ORACLE: both recover the payload"""
    rng = np.random.default_rng(808)
    tbs, Qm, G = 75376, 6, 90000
    data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
    got, rets = ctx.encode_tbs([(data, tbs, Qm, 0, G)])
    bits = np.unpackbits(got[0])[:G]
    llr = ((2 * bits.astype(np.int16) - 1) * 50).astype(np.int16)
    sbp = port.softbuffer_new()
    rc, out, nit, avg, crc = port.decode_tb(sbp, tbs, Qm, 0, llr, 4)
    port.softbuffer_del(sbp)
    assert rc == 0 and (out[:tbs // 8] == data).all()
    t = b.make_tbs(1)
    o = np.zeros(tbs // 8 + 22, np.uint8)
    t[0].e_bits, t[0].nof_e_bits, t[0].tbs, t[0].Qm, t[0].rv, t[0].data = llr.ctypes.data, G, tbs, Qm, 0, o.ctypes.data
    ctx.decode_tbs(t, False, 4)
    assert t[0].ret == 0 and (o[:tbs // 8] == data).all()


def test_front_end_and_encoder_reject_invalid_descriptors(ctx):
    sym = np.zeros(16, np.complex64)
    out = np.zeros(128, np.int16)
    dm = b.make_demods(1)
    dm[0].symbols, dm[0].nof_symbols, dm[0].mod, dm[0].scramble_bytes, dm[0].e_bits = sym.ctypes.data, 16, 7, None, out.ctypes.data
    with pytest.raises(b.B200Error):
        ctx.demod_descramble_raw(dm, False, 0)  # unknown modulation
    dm[0].mod, dm[0].nof_symbols = 2, 0
    with pytest.raises(b.B200Error):
        ctx.demod_descramble_raw(dm, False, 0)  # empty codeword
    # encoder: per-block return codes, the batch itself succeeds
    data = np.zeros(1000 // 8, np.uint8)
    _, rets = ctx.encode_tbs([(data, 1000, 2, 5, 2880), (data, 1000, 0, 0, 2880), (data, 1000, 2, 0, 2880)])
    assert rets == [-2, -2, 0]
    assert ctx.encode_tbs([]) == ([], [])


# ----------------------------------------------------------------------------------------- PUSCH pre-steps
def _ul_check(port, ctx, blocks):
    got = ctx.ulsch_deinterleave(blocks)
    for i, (q, Qm, nsym, qa, qr, qc) in enumerate(blocks):
        rc, g, ack, ri, _ = port.ulsch_deinterleave(q, Qm, nsym, qa, qr)
        assert rc == 0
        n = (len(q) // Qm - qr) * Qm
        assert (got[i][0] == g[:n]).all(), (i, Qm, nsym, qa, qr, qc, int(np.argmax(got[i][0] != g[:n])))
        assert (got[i][1] == ack).all() and (got[i][2] == ri).all() and (got[i][3] == g[:qc * Qm]).all(), (i, Qm, nsym, qa, qr, qc)
    return got


def test_ulsch_deinterleave(port, ctx):
    """k_ulsch_deinterleave against the oracle, many transport blocks per call: every Qm, normal / extended CP with and without
    SRS, matrices from one row to 1320 rows (110 PRB), no UCI, ACK only, RI only, both, up to the four-per-row maximum, CQI
    regions that do and do not cover the clobbered g_bits[0]; full-range int16 inputs"""
    rng = np.random.default_rng(3621)
    blocks = []
    for Qm in (2, 4, 6, 8):
        for nsym in (12, 11, 10, 9):
            for rows in (1, 2, 3, 12, 63, 64, 65, 300, 1200, 1320):
                if rows > 300 and (Qm == 8 or nsym in (11, 9)):
                    continue
                H = rows * nsym
                for qa, qr in ((0, 0), (1, 0), (0, 1), (3, 2), (4 * rows, 4 * rows), (min(4 * rows, 37), min(4 * rows, 5)), (min(4 * rows, 6), min(4 * rows, 41))):
                    qc = int(rng.integers(0, min(H - qr, 200) + 1)) if (rows + qa) % 2 else 0
                    blocks.append((rng.integers(-32768, 32768, H * Qm).astype(np.int16), Qm, nsym, qa, qr, qc))
    for nsym in (1, 7, 14):  # no UCI columns needed
        blocks.append((rng.integers(-32768, 32768, nsym * 50 * 4).astype(np.int16), 4, nsym, 0, 0, 3))
    assert len(blocks) > 500
    _ul_check(port, ctx, blocks[:1])
    _ul_check(port, ctx, blocks)
    # without the UCI outputs the call is a pure de-interleave
    got = ctx.ulsch_deinterleave(blocks[5:40], uci=False)
    for (q, Qm, nsym, qa, qr, qc), g in zip(blocks[5:40], got):
        assert (g[0] == port.ulsch_deinterleave(q, Qm, nsym, qa, qr)[1][:len(g[0])]).all()


def test_pusch_golden_on_device(ctx):
    """the fixtures recorded from the UNMODIFIED srslte_ulsch_decode (tests/golden/ulsch.npz): device de-interleave + device
    decode_tb reproduce its g_bits (digest), CQI LLRs, return code and decoded bytes"""
    import hashlib
    from util import UL_GRANTS
    u = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ulsch.npz"))
    dig = lambda a: np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], np.uint64)[0]
    blocks = []
    for i, grant in enumerate(UL_GRANTS[:10]):
        qa, qr, qc = u["qprime%d" % i].tolist()
        blocks.append((u["llr%d" % i], grant[1], grant[3], qa, qr, qc))
    got = ctx.ulsch_deinterleave(blocks)
    t = b.make_tbs(len(blocks))
    outs = []
    for i, (grant, (q, Qm, nsym, qa, qr, qc)) in enumerate(zip(UL_GRANTS, blocks)):
        g = got[i][0]
        # (the fixture covers the reference's whole g_bits array: its last Q'_ri * Qm elements are never written, fill 777)
        assert dig(np.concatenate([g[qc * Qm:], np.full(qr * Qm, 777, np.int16)])) == u["g_data%d" % i], i
        front = got[i][3].copy()
        if grant[6] == 1 and len(front) > 32:  # the reference's short-CQI decoder folds the copies in place (uci.c:379-385)
            for k in range(1, len(front) // 32):
                front[:32] += front[32 * k:32 * k + 32]
            k = len(front) // 32
            front[:len(front) % 32] += front[32 * k:]
        assert (front == u["g_front%d" % i]).all(), i
        e = np.ascontiguousarray(g[qc * Qm:])
        o = np.zeros(grant[0] // 8 + 22, np.uint8)
        outs.append((e, o))
        t[i].e_bits, t[i].nof_e_bits, t[i].tbs, t[i].Qm, t[i].rv, t[i].data = e.ctypes.data, len(e), grant[0], Qm, 0, o.ctypes.data
    ctx.decode_tbs(t, False, 8)
    for i, grant in enumerate(UL_GRANTS[:10]):
        assert t[i].ret == int(u["rc%d" % i][0]), i
        assert (outs[i][1][:grant[0] // 8] == u["data%d" % i]).all(), i


def test_pusch_symbols_to_transport_block_on_device(port, ctx):
    """the receive chain of pusch.c / srslte_ulsch_decode on the device for a 100-PRB 64QAM subframe carrying ACK, RI and CQI:
    symbols -> soft demodulation + descrambling -> UCI extraction + de-interleaving -> decode_tb, LLRs never leaving the GPU,
    against the same chain through the oracle"""
    from srsran_b200 import synth
    from util import ul_interleave_tx
    rng = np.random.default_rng(1105)
    tbs, mod, Qm, L_prb, nsym = 75376, 3, 6, 100, 12
    rows, H = L_prb * 12, L_prb * 12 * nsym
    qa, qr, qc = 36, 20, 57
    G = (H - qr - qc) * Qm
    data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
    g_tx = np.concatenate([rng.integers(0, 2, qc * Qm, dtype=np.uint8), port.encode_tb(tbs, Qm, 0, G, data)])
    q_tx = ul_interleave_tx(port, rng, g_tx, Qm, rows, nsym, qa, qr)
    scr = port.sequence_bytes((0x46 << 14) + (5 << 9) + 1, H * Qm)
    sym = synth.lte_modulate(q_tx ^ np.unpackbits(scr)[:H * Qm], mod)
    sym = (sym + 0.07 * (rng.standard_normal(H) + 1j * rng.standard_normal(H))).astype(np.complex64)
    # oracle chain
    q_llr = port.descramble(scr, port.demod(mod, sym, np.int16))
    rc0, g, ack, ri, _ = port.ulsch_deinterleave(q_llr, Qm, nsym, qa, qr)
    sbp = port.softbuffer_new()
    rc, want, nit, avg, crc = port.decode_tb(sbp, tbs, Qm, 0, g[qc * Qm:qc * Qm + G].copy(), 8)
    port.softbuffer_del(sbp)
    assert rc0 == 0 and rc == 0 and (want[:tbs // 8] == data).all()
    # device chain
    d_q, d_g, d_out = ctx.device_alloc(H * Qm * 2 + 64), ctx.device_alloc(H * Qm * 2 + 64), ctx.device_alloc(tbs // 8 + 64)
    dm = b.make_demods(1)
    dm[0].symbols, dm[0].nof_symbols, dm[0].mod, dm[0].scramble_bytes, dm[0].e_bits = sym.ctypes.data, H, mod, scr.ctypes.data, d_q
    ctx.demod_descramble_raw(dm, False, b.OUT_DEVICE)
    ul = b.make_ulschs(1)
    a_l, r_l, c_l = np.zeros(qa * Qm, np.int16), np.zeros(qr * Qm, np.int16), np.zeros(qc * Qm, np.int16)
    ul[0].q_bits, ul[0].Qm, ul[0].H_prime_total, ul[0].N_pusch_symbs, ul[0].g_bits = d_q, Qm, H, nsym, d_g
    ul[0].Q_prime_ack, ul[0].Q_prime_ri, ul[0].Q_prime_cqi = qa, qr, qc
    ul[0].ack_llr, ul[0].ri_llr, ul[0].cqi_llr = a_l.ctypes.data, r_l.ctypes.data, c_l.ctypes.data
    ctx.ulsch_deinterleave_raw(ul, b.IN_DEVICE | b.OUT_DEVICE)
    assert (a_l == ack).all() and (r_l == ri).all() and (c_l == g[:qc * Qm]).all()
    # SRSLTE_B200_UCI_DEFERRED: the same call only enqueues; the UCI LLRs arrive with the next wait() on the context
    a_l[:], r_l[:], c_l[:] = 0, 0, 0
    ctx.ulsch_deinterleave_raw(ul, b.IN_DEVICE | b.OUT_DEVICE | b.UCI_DEFERRED)
    ctx.wait()
    assert (a_l == ack).all() and (r_l == ri).all() and (c_l == g[:qc * Qm]).all()
    a_l[:], r_l[:], c_l[:] = 0, 0, 0
    ctx.ulsch_deinterleave_raw(ul, b.IN_DEVICE | b.OUT_DEVICE | b.UCI_DEFERRED)
    t = b.make_tbs(1)
    t[0].e_bits, t[0].nof_e_bits, t[0].tbs, t[0].Qm, t[0].rv, t[0].softbuffer, t[0].data = d_g + qc * Qm * 2, G, tbs, Qm, 0, None, d_out
    ctx.decode_tbs(t, False, 8, flags=b.IN_DEVICE | b.OUT_DEVICE)
    out = np.zeros(tbs // 8 + 6, np.uint8)
    ctx.d2h(out, d_out)
    for p in (d_q, d_g, d_out):
        ctx.device_free(p)
    assert (a_l == ack).all() and (r_l == ri).all() and (c_l == g[:qc * Qm]).all()  # (filled by decode_tbs' wait)
    assert t[0].ret == 0 and (out == want[:tbs // 8 + 6]).all() and (out[:tbs // 8] == data).all()
    assert list(t[0].cb_noi[:13]) == nit[:13].tolist()


def test_ulsch_rejects_invalid_descriptors(ctx):
    q, g = np.zeros(12 * 8 * 2, np.int16), np.zeros(12 * 8 * 2, np.int16)
    def call(Qm=2, H=96, nsym=12, qa=0, qr=0, qc=0, qp=q.ctypes.data, gp=g.ctypes.data):
        ul = b.make_ulschs(1)
        ul[0].q_bits, ul[0].Qm, ul[0].H_prime_total, ul[0].N_pusch_symbs, ul[0].g_bits = qp, Qm, H, nsym, gp
        ul[0].Q_prime_ack, ul[0].Q_prime_ri, ul[0].Q_prime_cqi = qa, qr, qc
        ctx.ulsch_deinterleave_raw(ul, 0)
    call()
    for bad in (dict(Qm=3), dict(Qm=0), dict(Qm=10), dict(H=95), dict(nsym=0), dict(nsym=15, H=90), dict(qa=33), dict(qr=33), dict(qr=8, qc=89), dict(qp=None),
                dict(gp=None), dict(nsym=8, qr=1), dict(nsym=6, H=96, qa=1)):
        with pytest.raises(b.B200Error):
            call(**bad)
    assert ctx.ulsch_deinterleave([]) == []


# ----------------------------------------------------------------------------------------- BASELINE config 5 / SURVEY 8e: pooled cells
def test_pooled_cells_with_harq(port):
    """srsran_b200.pool.CellPool: downlink-shaped and uplink subframes of several cells over a run of TTIs, cells sharded over
    two ranks (both on this GPU), HARQ processes with device-resident soft buffers -- failed transport blocks are
    retransmitted with the next redundancy version into the same process -- against the oracle walking the same job
    sequence with its own soft buffers"""
    from srsran_b200.pool import CellPool, Job, owner_of, partition
    from util import ul_interleave_tx
    rng = np.random.default_rng(2005)
    world = 2
    pools = [CellPool(rank=r, world=world, device=0) for r in range(world)]
    grants = {  # per cell: (dl: tbs, Qm, G, sigma), (ul: tbs, Qm, L_prb, nsym, q_prime, sigma)
        0: ((15264, 4, 20000, 0.74), (9912, 4, 25, 12, (23, 8, 10), 0.62)),
        1: ((75376, 6, 90000, 0.93), (5160, 2, 25, 12, (12, 4, 5), 1.05)),
        2: ((2216, 2, 7000, 1.45), (75376, 6, 100, 12, (36, 20, 57), 0.83)),
        3: ((31704, 6, 40000, 0.80), (3624, 2, 25, 10, (28, 3, 5), 0.95)),
    }
    rvs = (0, 2, 3, 1)
    state = {}   # (cell, kind, pid) -> [data, transmissions so far]
    orc_sb = {}
    n_retx = n_fail_first = n_jobs = 0
    for tti in range(10):
        jobs = []
        for cell, (gdl, gul) in grants.items():
            for kind in ("dl", "ul"):
                pid = tti % 4
                key = (cell, kind, pid)
                st = state.get(key)
                tbs = gdl[0] if kind == "dl" else gul[0]
                if st is None:
                    st = state[key] = [rng.integers(0, 256, tbs // 8, dtype=np.uint8), 0]
                data, ntx = st
                rv = rvs[ntx]
                if kind == "dl":
                    _, Qm, G, sigma = gdl
                    llr = bpsk_awgn_llr(rng, port.encode_tb(tbs, Qm, rv, G, data), 100, sigma, np.int16)
                    jobs.append(Job(cell, tti, "dl", pid, rv, ntx == 0, tbs, Qm, llr))
                else:
                    _, Qm, L_prb, nsym, (qa, qr, qc), sigma = gul
                    rows = L_prb * 12
                    G = (rows * nsym - qr - qc) * Qm
                    g_tx = np.concatenate([rng.integers(0, 2, qc * Qm, dtype=np.uint8), port.encode_tb(tbs, Qm, rv, G, data)])
                    llr = bpsk_awgn_llr(rng, ul_interleave_tx(port, rng, g_tx, Qm, rows, nsym, qa, qr), 100, sigma, np.int16)
                    jobs.append(Job(cell, tti, "ul", pid, rv, ntx == 0, tbs, Qm, llr, nsym, (qa, qr, qc)))
        # every rank decodes its own cells; together they cover the batch exactly once
        parts = [partition(jobs, r, world) for r in range(world)]
        assert sorted(id(j) for p in parts for j in p) == sorted(id(j) for j in jobs)
        results = {}
        for r in range(world):
            assert all(owner_of(j.cell, world) == r for j in parts[r])
            for j, res in zip(parts[r], pools[r].decode(parts[r])):
                results[id(j)] = res
        for j in jobs:
            key = (j.cell, j.kind, j.pid)
            res = results[id(j)]
            if key not in orc_sb:
                orc_sb[key] = port.softbuffer_new()
            if j.new_data:
                port.softbuffer_reset(orc_sb[key])
            e = j.llr
            if j.kind == "ul":
                qa, qr, qc = j.q_prime
                rc0, g, ack, ri, _ = port.ulsch_deinterleave(j.llr, j.Qm, j.n_pusch_symbs, qa, qr)
                assert rc0 == 0 and (res.ack_llr == ack).all() and (res.ri_llr == ri).all() and (res.cqi_llr == g[:qc * j.Qm]).all()
                e = g[qc * j.Qm:(len(j.llr) // j.Qm - qr) * j.Qm].copy()
            rc, want, nit, avg, crc = port.decode_tb(orc_sb[key], j.tbs, j.Qm, j.rv, e, 8)
            C_ = len(res.cb_noi)
            assert res.ret == rc, (tti, key, res.ret, rc)
            assert (res.data[:j.tbs // 8 + 3] == want[:j.tbs // 8 + 3]).all(), (tti, key)
            assert res.cb_noi == nit[:C_].tolist() and res.cb_crc == crc[:C_].tolist(), (tti, key)
            assert abs(res.avg_iterations - avg) < 1e-6
            n_jobs += 1
            st = state[key]
            if rc == 0:
                assert (res.data[:j.tbs // 8] == st[0]).all()
                del state[key]
            else:
                n_fail_first += st[1] == 0
                st[1] += 1
                n_retx += 1
                if st[1] == 4:
                    del state[key]
    assert n_jobs == 80 and n_fail_first >= 8 and n_retx >= 8, (n_jobs, n_fail_first, n_retx)  # the run really exercises HARQ combining
    with pytest.raises(b.B200Error):
        pools[0].decode([Job(1, 0, "dl", 0, 0, True, 2216, 2, np.zeros(7000, np.int16))])  # cell 1 belongs to rank 1
    for sb in orc_sb.values():
        port.softbuffer_del(sb)
    for p in pools:
        p.close()


def test_ulsch_deinterleave_device_buffers_stay_in_bounds(port, ctx):
    """device-resident in and out (SRSLTE_B200_IN_DEVICE | OUT_DEVICE) with sentinel-filled guard bands around every g_bits
    array and after the unwritten RI tail: nothing outside [0, (H' - Q'_ri) * Qm) is touched"""
    rng = np.random.default_rng(77)
    geo = [(6, 12, 1200, 36, 20, 57), (6, 12, 1200, 0, 0, 0), (2, 10, 300, 28, 3, 5), (4, 11, 36, 7, 2, 0), (4, 12, 129, 516, 516, 40), (2, 9, 1, 4, 4, 0),
           (8, 12, 65, 3, 9, 700), (6, 12, 128, 1, 1, 1), (2, 12, 127, 0, 5, 0)]
    guard = 256  # int16 elements
    ul = b.make_ulschs(len(geo))
    bufs = []
    for i, (Qm, nsym, rows, qa, qr, qc) in enumerate(geo):
        n = rows * nsym * Qm
        q = rng.integers(-32768, 32768, n).astype(np.int16)
        d_q = ctx.device_alloc(n * 2 + 64)
        d_g = ctx.device_alloc((n + 2 * guard) * 2)
        ctx.h2d(d_q, q)
        ctx.h2d(d_g, np.full(n + 2 * guard, 31111, np.int16))
        ul[i].q_bits, ul[i].Qm, ul[i].H_prime_total, ul[i].N_pusch_symbs, ul[i].g_bits = d_q, Qm, rows * nsym, nsym, d_g + guard * 2
        ul[i].Q_prime_ack, ul[i].Q_prime_ri, ul[i].Q_prime_cqi = qa, qr, qc
        bufs.append((q, d_q, d_g, n))
    ctx.ulsch_deinterleave_raw(ul, b.IN_DEVICE | b.OUT_DEVICE)
    for (Qm, nsym, rows, qa, qr, qc), (q, d_q, d_g, n) in zip(geo, bufs):
        out = np.zeros(n + 2 * guard, np.int16)
        ctx.d2h(out, d_g)  # (memcpy_d2h drains the context's stream first)
        rc, g, _, _, _ = port.ulsch_deinterleave(q, Qm, nsym, qa, qr, g_fill=31111)
        assert rc == 0 and (out[guard:guard + n] == g).all(), (Qm, nsym, rows, qa, qr, qc)
        assert (out[:guard] == 31111).all() and (out[guard + n:] == 31111).all()
        ctx.device_free(d_q)
        ctx.device_free(d_g)


# ----------------------------------------------------------------------------------------- both MAP kernels on the same inputs
def test_latency_option_both_kernels(port, ctx):
    """Small batches run the latency-shaped k_map_lat by default, so the throughput kernel k_map_f16 would lose its
    small-case coverage: every case here is decoded twice, with the option "latency" off (k_map_f16) and on (k_map_lat), and
    both must equal the oracle -- code-block batches of every 16-lane size class incl. partial top tiles (K = 5824, 816,
    1008) and saturating amplitudes (exact replay), and transport blocks with CRC early stop and HARQ combining."""
    rng = np.random.default_rng(4848)
    cb_cases = []
    for K, amp, nit in ((6144, 300, 4), (6144, 30000, 5), (5824, 2000, 6), (816, 5000, 4), (1008, 700, 5), (2048, 100, 4), (4160, 900, 3), (3136, 32767, 4)):
        llr = np.stack([random_llr(rng, 3 * (K + 32) + 12, amp, np.int16) for _ in range(5)])
        hp = port.tdec_new(TDEC_AUTO, False)
        want = [port.tdec_run_all(hp, llr[i], nit, K)[1] for i in range(5)]
        port.tdec_del(hp)
        cb_cases.append((K, nit, llr, want))
    tb_cases = []
    for tbs, Qm, G, sigma in ((75376, 6, 90000, 0.46), (75376, 6, 90000, 0.95), (15264, 4, 20000, 0.75), (31704, 6, 40000, 0.5)):
        data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        tb_cases.append((tbs, Qm, G, [(rv, _tb_inputs(port, rng, tbs, Qm, G, rv, np.int16, 100, sigma, data)[1]) for rv in (0, 2)]))
    try:
        for lat in (0, 1):
            ctx.set_option("latency", lat)
            for K, nit, llr, want in cb_cases:
                got = ctx.tdec_batch(llr, K, nit, input_sb=True)
                for i in range(5):
                    assert (got[i] == want[i]).all(), (lat, K, i)
            for tbs, Qm, G, txs in tb_cases:
                sbp, sbg = port.softbuffer_new(), ctx.softbuffer_create()
                for rv, llr in txs:
                    _decode_both(port, ctx, tbs, Qm, rv, llr, 8, sbp, sbg)
                port.softbuffer_del(sbp)
                ctx.softbuffer_free(sbg)
    finally:
        ctx.set_option("latency", 1)
    with pytest.raises(b.B200Error):
        ctx.set_option("no-such-option", 1)


# ----------------------------------------------------------------------------------------- the fused persistent kernel at the benchmarked scale
def _scale_batch(rng, port, K, n_base, amps, dtype, sb):
    """n_base distinct AWGN code words (amplitude cycles over `amps`, so some blocks trip the Fast16 range monitor)"""
    N = (lanes16(K) if dtype == np.int16 else lanes8(K)) if sb else 0
    rows = []
    for i in range(n_base):
        bits = rng.integers(0, 2, K, dtype=np.uint8)
        llr = bpsk_awgn_llr(rng, port.tcod_encode(bits), amps[i % len(amps)], 0.75 + 0.5 * (i % 7) / 6, dtype)
        rows.append(std_to_sb(llr, K, N) if N else llr)
    return np.stack(rows)


def _spread(n, n_base):
    """which distinct block sits at batch position i: every base block appears in many groups, warps and waves"""
    i = np.arange(n)
    return (i * 7 + i // n_base) % n_base


@pytest.mark.parametrize("K,dtype,ncb,nit,amps,sb", [
    (6144, np.int16, 8000, 4, (100, 100, 100, 260, 2500), False),   # the benchmark's shape: several waves of 12-warp CTAs, some blocks parked
    (6144, np.int16, 4100, 8, (100, 180, 300), True),
    (5824, np.int16, 8000, 5, (100, 100, 400), True),               # partial top tile (W = 364)
    (6144, np.int8, 4000, 6, (20, 45, 127), True),                  # Sat8, 32 lanes
    (2048, np.int8, 5000, 5, (25, 127), True),                      # Sat8, 16 lanes
    (512, np.int16, 6000, 6, (100, 900), True),                     # 8-lane class, 8 blocks per warp
    (1008, np.int16, 3000, 7, (100, 3000), False),                  # W = 63: odd, every flush of the decision bits unaligned
])
def test_fused_kernel_at_scale(port, K, dtype, ncb, nit, amps, sb):
    """k_map_fused (NOT the latency-shaped kernel small batches get) at the size the benchmark runs it: thousands of blocks
    of one size on two engines that decode concurrently; every block of the batch is compared with the oracle's result for
    its distinct input (duplicates at different slots, warps and waves must decode identically)."""
    rng = np.random.default_rng(K + ncb + nit)
    n_base = 56
    base = _scale_batch(rng, port, K, n_base, amps, dtype, sb)
    hp = port.tdec_new(TDEC_AUTO, not sb)
    want = np.stack([port.tdec_run_all(hp, base[i], nit, K)[1] for i in range(n_base)])
    port.tdec_del(hp)
    idx = _spread(ncb, n_base)
    batch = np.ascontiguousarray(base[idx])
    ctxs = [b.Context(0), b.Context(0)]
    outs = [np.zeros((ncb, K // 8), np.uint8) for _ in ctxs]
    try:
        for c, o in zip(ctxs, outs):  # both submitted before either is waited for
            c.tdec_batch_submit(batch.ctypes.data, o.ctypes.data, K, ncb, batch.shape[1], 16 if dtype == np.int16 else 8, nit, input_sb=sb)
        for c in ctxs:
            c.wait()
            assert c.last_half_iterations() == ncb * nit
            assert c.last_map_launches() <= 2  # one persistent launch (+ the exact-arithmetic launch for parked blocks)
        for o in outs:
            bad = np.nonzero((o != want[idx]).any(axis=1))[0]
            assert len(bad) == 0, "K=%d: %d of %d blocks differ from the oracle, first at %d (base %d)" % (K, len(bad), ncb, bad[0], idx[bad[0]])
        if dtype == np.int16 and max(amps) > 1000:
            assert 0 < ctxs[0].last_replayed() < ncb * nit  # some blocks really took the parked / exact path
    finally:
        for c in ctxs:
            c.close()


def test_fused_kernel_transport_blocks_at_scale(port):
    """the transport-block path through k_map_fused: 360 blocks of 75376 bits (13 x K=5824) + 300 of 97896 bits in int8 on a
    second engine at the same time, noise levels spread so that code blocks stop after 1..8 half-iterations or never; return
    codes, bytes, per-block CRC flags and half-iteration counts of EVERY transport block against the oracle"""
    rng = np.random.default_rng(2602)
    jobs = []
    for tbs, Qm, G, dtype, amp, sigmas, ntb in ((75376, 6, 90000, np.int16, 100, (0.3, 0.44, 0.5, 0.56, 0.62, 0.9), 360),
                                                (97896, 8, 115200, np.int8, 20, (0.3, 0.40, 0.47, 0.55), 300)):
        base, want = [], []
        for i, sg in enumerate(sigmas * 2):
            _, llr = _tb_inputs(port, rng, tbs, Qm, G, 0, dtype, amp, sg)
            sbp = port.softbuffer_new()
            want.append(port.decode_tb(sbp, tbs, Qm, 0, llr, 8))
            port.softbuffer_del(sbp)
            base.append(llr)
        idx = _spread(ntb, len(base))
        t = b.make_tbs(ntb)
        outs = np.zeros((ntb, tbs // 8 + 22), np.uint8)
        for i in range(ntb):
            t[i].e_bits, t[i].nof_e_bits, t[i].tbs, t[i].Qm, t[i].rv, t[i].data = base[idx[i]].ctypes.data, G, tbs, Qm, 0, outs[i].ctypes.data
        jobs.append((b.Context(0), t, dtype == np.int8, tbs, idx, base, want, outs))
    try:
        for c, t, is8, *_ in jobs:
            c.decode_tbs(t, is8, 8, submit_only=True)
        for c, t, is8, tbs, idx, base, want, outs in jobs:
            c.wait()
            seen = set()
            for i in range(len(t)):
                rc, d, nit, avg, crc = want[idx[i]]
                C_ = t[i].nof_cb
                nb = tbs // 8 + 6
                assert t[i].ret == rc, (tbs, i)
                assert (outs[i][:nb] == d[:nb]).all(), (tbs, i)
                assert list(t[i].cb_noi[:C_]) == nit[:C_].tolist(), (tbs, i)
                assert list(t[i].cb_crc[:C_]) == crc[:C_].tolist(), (tbs, i)
                seen.update(nit[:C_].tolist())
            assert len(seen) >= 4 and c.last_map_launches() <= 2
    finally:
        for j in jobs:
            j[0].close()


# ----------------------------------------------------------------------------------------- short blocks: the fused generic kernel
@pytest.mark.gpu
@pytest.mark.parametrize("K,ncb,nit,amps", [(40, 1001, 6, (100, 3000)), (400, 777, 5, (100, 20000, 32767)), (208, 2, 3, (300,)), (104, 1, 8, (50,)),
                                            (320, 1500, 1, (600,)), (360, 901, 2, (150, 9000))])
def test_generic_decoder_fused_kernel_at_scale(port, K, ncb, nit, amps):
    """k_gen_fused (K <= 400 under SRSLTE_TDEC_AUTO): batches of many CTAs with an odd block at the end (a pair with one half
    empty), amplitudes up to the int16 wrap cases of turbodecoder_gen.c; decided bytes of EVERY block against the oracle,
    the extrinsic planes of some against the oracle's arrays, and everything against the per-half-iteration kernel"""
    rng = np.random.default_rng(K * 7 + ncb)
    base = [bpsk_awgn_llr(rng, port.tcod_encode(rng.integers(0, 2, K, dtype=np.uint8)), amps[i % len(amps)], 0.7 + 0.15 * (i % 4), np.int16) for i in range(7)]
    idx = _spread(ncb, len(base))
    batch = np.ascontiguousarray(np.stack([base[j] for j in idx]))
    hs, want = [], []
    for x in base:
        h = port.tdec_new(TDEC_AUTO, False)
        rc, w = port.tdec_run_all(h, x, nit, K)
        assert rc == 0
        hs.append(h)
        want.append(w)
    c = b.Context(0)
    try:
        got = c.tdec_batch(batch, K, nit)
        assert c.last_map_launches() == 1
        for i in range(ncb):
            assert (got[i] == want[idx[i]]).all(), (K, i)
        for cb in sorted({0, 1, ncb // 2, ncb - 2, ncb - 1} & set(range(ncb))):
            plane, w = _expected_planes(port, hs[idx[cb]], K, 0, nit, 16)
            g = c.debug_read_plane(cb, plane, K)
            assert (g == w).all(), "K=%d n=%d block %d: %d LLRs differ" % (K, nit, cb, int((g != w).sum()))
        c.set_option("gen_fused", 0)
        old = c.tdec_batch(batch, K, nit)
        assert (old == got).all()
    finally:
        c.close()
        for h in hs:
            port.tdec_del(h)


@pytest.mark.gpu
def test_generic_decoder_fused_kernel_transport_blocks(port):
    """single-block transport blocks of every generic-decoder size (40..400) in one batch, 9 of each with noise levels that
    make blocks stop after 1..8 half-iterations or fail: return codes, bytes, CRC flags and half-iteration counts of every
    block against the oracle; pairs whose halves stop at different times are the rule here"""
    rng = np.random.default_rng(4004)
    sizes = [K for K in all_K() if K <= 400]
    cases, base = [], {}
    for K in sizes:
        tbs = K - 24
        G = 2 * ((3 * K + 12) // 2 + (K // 8 % 3) * 30)
        for j, sg in enumerate((0.5, 0.8, 1.1)):
            _, llr = _tb_inputs(port, rng, tbs, 2, G, 0, np.int16, 100, sg)
            sbp = port.softbuffer_new()
            base[(K, j)] = (llr, G, port.decode_tb(sbp, tbs, 2, 0, llr, 8))
            port.softbuffer_del(sbp)
        for r in range(9):
            cases.append((K, (r * 2 + K // 8) % 3))
    n = len(cases)
    t = b.make_tbs(n)
    outs = np.zeros((n, 400 // 8 + 22), np.uint8)
    for i, (K, j) in enumerate(cases):
        llr, G, _ = base[(K, j)]
        t[i].e_bits, t[i].nof_e_bits, t[i].tbs, t[i].Qm, t[i].rv, t[i].data = llr.ctypes.data, G, K - 24, 2, 0, outs[i].ctypes.data
    c = b.Context(0)
    try:
        c.decode_tbs(t, False, 8)
        assert c.last_map_launches() == 1
        seen = set()
        for i, (K, j) in enumerate(cases):
            rc, d, nit, avg, crc = base[(K, j)][2]
            nb = (K - 24) // 8 + 3
            assert t[i].ret == rc, (K, i)
            assert (outs[i][:nb] == d[:nb]).all(), (K, i)
            assert t[i].cb_noi[0] == nit[0] and t[i].cb_crc[0] == crc[0], (K, i)
            seen.add(int(nit[0]))
        assert len(seen) >= 4
    finally:
        c.close()


# ----------------------------------------------------------------------------------------- intermediate LLRs, not only decided bytes
def _expected_planes(port, h, K, N, n_done, bits):
    """what the engine's a-priori (plane 2) / decoder-2 input (plane 3) planes must hold after n_done half-iterations, from the
    oracle's arrays: the glue subtractions of turbodecoder_iter.h:107-109 / 117-119 are fused into the kernels' epilogues"""
    app1 = port.tdec_get_llr(h, 0, K).astype(np.int64)
    ext1 = port.tdec_get_llr(h, 2, K).astype(np.int64)
    fwd, rev = b.qpp_table(K, N)

    def sub(a, c):
        d = a - c
        if bits == 16:
            return ((d + 32768) % 65536 - 32768).astype(np.int16)  # srslte_vec_sub_sss wraps
        sat = np.clip(d, -128, 127)
        wrap = (d + 128) % 256 - 128
        idx = np.arange(K)
        return np.where(idx < (K // 32) * 32, sat, wrap).astype(np.int16)  # srslte_vec_sub_bbb, AVX2 build
    if n_done % 2 == 1:
        e = sub(ext1, app1) if n_done > 1 else ext1.astype(np.int16)  # iter.h:117-119, then app2[rev[i]] = ext1[i]
        want = np.zeros(K, np.int16)
        want[rev] = e
        return 3, want
    return 2, sub(app1, ext1)  # iter.h:107-109 (the subtraction the next DEC1 would perform)


@pytest.mark.parametrize("K,dtype,ncb,amp", [(6144, np.int16, 700, 100), (6144, np.int16, 9, 100), (5824, np.int16, 640, 150), (1008, np.int16, 1300, 120),
                                              (512, np.int16, 1400, 100), (6144, np.int8, 400, 25), (2048, np.int8, 900, 30)])
def test_intermediate_llrs_match_the_oracle(port, ctx, K, dtype, ncb, amp):
    """LLR-for-LLR: after n = 1..5 half-iterations the extrinsic planes in the engine's workspace equal what the oracle's
    arrays say, for blocks at the start, the middle and the end of large batches (k_map_fused) and of a small one (k_map_lat)"""
    rng = np.random.default_rng(K + ncb)
    bits = 16 if dtype == np.int16 else 8
    N = lanes16(K) if bits == 16 else lanes8(K)
    base = [std_to_sb(bpsk_awgn_llr(rng, port.tcod_encode(rng.integers(0, 2, K, dtype=np.uint8)), amp, 0.8 + 0.1 * i, dtype), K, N) for i in range(3)]
    batch = np.ascontiguousarray(np.stack([base[i % 3] for i in range(ncb)]))
    probe = sorted({0, 1, 2, ncb // 2, ncb - 3, ncb - 2, ncb - 1})
    for n in (1, 2, 3, 4, 5):
        ctx.tdec_batch(batch, K, n, input_sb=True)
        hs = []
        for i in range(3):
            h = port.tdec_new(TDEC_AUTO, False)
            port.tdec_run_all(h, base[i], n, K)
            hs.append(h)
        for cb in probe:
            plane, want = _expected_planes(port, hs[cb % 3], K, N, n, bits)
            got = ctx.debug_read_plane(cb, plane, K)
            assert (got == want).all(), "K=%d bits=%d n=%d block %d: %d LLRs differ" % (K, bits, n, cb, int((got != want).sum()))
        for h in hs:
            port.tdec_del(h)


@pytest.mark.parametrize("K,ncb,amps", [(6144, 13, (100,)), (5824, 13, (100, 900)), (6144, 26, (300, 3000, 12000)), (4096, 9, (100, 20000)), (3136, 5, (700,)),
                                        (3072, 32, (100, 5000)), (6144, 1, (32767,))])
def test_time_parallel_latency_kernels(port, ctx, K, ncb, amps):
    """A subframe or two of int16 blocks runs the time-parallel kernels of map_scan.cuh: the cooperative k_scan_fused (default),
    the per-half-iteration pair k_scan_mat + k_scan_out (option scan_launch) -- both against k_map_lat (option scan off) and the
    oracle: decided bytes after 1..7 half-iterations and the extrinsic planes LLR for LLR.  Amplitudes are mixed inside a
    group so that the range monitor parks some blocks and not others, at different half-iterations (exact kernel takes over)."""
    rng = np.random.default_rng(K + ncb)
    N = lanes16(K)
    base = [std_to_sb(bpsk_awgn_llr(rng, port.tcod_encode(rng.integers(0, 2, K, dtype=np.uint8)), amps[i % len(amps)], 0.7 + 0.1 * (i % 3), np.int16), K, N)
            for i in range(min(ncb, 5))]
    batch = np.ascontiguousarray(np.stack([base[i % len(base)] for i in range(ncb)]))
    try:
        for n in (1, 2, 4, 5, 7):
            hs, want = [], []
            for b0 in base:
                h = port.tdec_new(TDEC_AUTO, False)
                want.append(port.tdec_run_all(h, b0, n, K)[1])
                hs.append(h)
            for scan, launch, fused in ((0, 0, 0), (1, 0, 1), (1, 1, 0)):
                ctx.set_option("scan", scan)
                ctx.set_option("scan_launch", launch)
                ctx.set_option("scan_fused", fused)
                got = ctx.tdec_batch(batch, K, n, input_sb=True)
                if scan and fused and ncb <= 32 and K >= 16 * 192:
                    assert ctx.last_map_launches() == 2, "k_scan_fused + the exact kernel for parked blocks"
                for i in range(ncb):
                    assert (got[i] == want[i % len(base)]).all(), (K, n, scan, launch, fused, i)
                for cb in sorted({0, ncb // 2, ncb - 1}):
                    plane, exp = _expected_planes(port, hs[cb % len(base)], K, N, n, 16)
                    llr = ctx.debug_read_plane(cb, plane, K)
                    assert (llr == exp).all(), "K=%d n=%d scan=%d/%d/%d block %d: %d LLRs differ" % (K, n, scan, launch, fused, cb, int((llr != exp).sum()))
            for h in hs:
                port.tdec_del(h)
    finally:
        ctx.set_option("scan", 1)
        ctx.set_option("scan_launch", 0)
        ctx.set_option("scan_fused", 1)


def test_int8_full_scale_every_windowed_size(port, ctx):
    """int8 LLRs over the whole +-127 range on EVERY size the int8 decoders take in lane layout (K >= 408: widened 8-lane,
    16-lane and 32-lane decoders), 4 half-iterations, decided bytes after the run; batches large enough for k_map_fused on the
    16- and 32-lane classes every few sizes, small ones (k_map_lat) in between"""
    rng = np.random.default_rng(8127)
    for n, K in enumerate(k for k in all_K() if k >= 408):
        N = lanes8(K)
        ncb = 1300 if n % 6 == 0 else 5
        base = [random_llr(rng, 3 * (K + 32) + 12, 127, np.int8) for _ in range(2)]
        batch = np.ascontiguousarray(np.stack([base[i % 2] for i in range(ncb)]))
        got = ctx.tdec_batch(batch, K, 4, input_sb=True)
        hp = port.tdec_new(TDEC_AUTO, False)
        want = [port.tdec_run_all(hp, base[i], 4, K)[1] for i in range(2)]
        port.tdec_del(hp)
        for i in range(ncb):
            assert (got[i] == want[i % 2]).all(), (K, N, i)


# ----------------------------------------------------------------------------------------- INTEGRATION.md option A, for real
def test_unmodified_sch_c_on_the_b200_library(port):
    """The reference's UNMODIFIED lib/src/phy/phch/sch.c -- srslte_dlsch_decode2 -> decode_tb -> decode_tb_cb with its per-code-
    block loop over srslte_rm_turbo_rx_lut[_8bit], srslte_tdec_new_cb / srslte_tdec_iteration[_8bit] and
    srslte_crc_checksum_byte, its softbuffer->buffer_f[] / data[] accesses and its HARQ bookkeeping -- compiled against
    include/srslte_b200/fec.h and linked against libsrslte_fec_b200.so in place of the reference's FEC objects
    (oracle/build_opt_a.sh).  Return codes, bytes, CRC flags and average iteration counts equal the oracle's, incl. a HARQ
    retransmission that combines in the HOST soft buffer the reference's loop keeps."""
    from oracle.bindings import OptA
    if not OptA.available():
        pytest.skip("oracle/_ref/libsch_on_b200.so not built (oracle/build_opt_a.sh needs /root/reference)")
    rng = np.random.default_rng(31337)
    for tbs, Qm, G, is8, amp, sigma, rvs in ((15264, 4, 20000, False, 200, 0.55, (0,)), (15264, 4, 20000, False, 200, 0.78, (0, 2, 3)),
                                              (6120, 2, 14400, True, 30, 0.8, (0,)), (75376, 6, 90000, False, 100, 0.46, (0,))):
        s = OptA(is8, 8)
        sb = port.softbuffer_new()
        s.reset_rx(tbs)
        data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        for rv in rvs:
            s.encode(tbs, Qm, G, 0, data)
            e = s.encode(tbs, Qm, G, rv, data)  # the reference's CPU encoder (CRC attach through the library's symbols)
            assert (e == port.encode_tb(tbs, Qm, rv, G, data)).all()
            llr = bpsk_awgn_llr(rng, e, amp, sigma, np.int8 if is8 else np.int16)
            rc, d, avg, crc = s.decode(tbs, Qm, rv, llr)
            rc_p, d_p, nit_p, avg_p, crc_p = port.decode_tb(sb, tbs, Qm, rv, llr, 8)
            C_ = port.cbsegm(tbs)[1]["C"]
            assert rc == rc_p, (tbs, rv, rc, rc_p)
            assert (d[:tbs // 8 + 3] == d_p[:tbs // 8 + 3]).all() and abs(avg - avg_p) < 1e-6, (tbs, rv)
            assert crc[:C_].tolist() == crc_p[:C_].tolist()
            if rc == 0:
                assert (d[:tbs // 8] == data).all()
                break
        s.close()
        port.softbuffer_del(sb)


@pytest.mark.parametrize("dtype", [np.int16, np.int8])
def test_demod_with_csi_correction(port, ctx, dtype):
    """the whole front end of pdsch.c:832-852 with csi_enable: soft demodulation -> csi_correction (pdsch.c:628-741) ->
    descrambling in one kernel, against the oracle's three functions in sequence (each pinned to the reference): every
    modulation, lengths around the SSE group sizes of all three steps, codewords with and without csi in one call"""
    rng = np.random.default_rng(741 + (dtype == np.int8))
    cws, want = [], []
    for mod in range(5):
        for n in (2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 33, 100, 1001, 15000):
            amp = (0.3, 1.0, 3.0, 50.0)[(n + mod) % 4]
            sym = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * amp).astype(np.complex64)
            nbits = n * b.MOD_BITS[mod]
            scr = port.sequence_bytes(int(rng.integers(1, 2 ** 31)), nbits) if (n + mod) % 3 else None
            csi = (rng.random(n).astype(np.float32) * 1.3 + 0.02) if (n + mod) % 4 else None
            cws.append((sym, mod, scr, csi))
            llr = port.demod(mod, sym, dtype)
            if csi is not None:
                llr = port.csi_correction(csi, llr, mod)
            want.append(port.descramble(scr, llr) if scr is not None else llr)
    got = ctx.demod_descramble(cws, dtype)
    for i, (g, w) in enumerate(zip(got, want)):
        assert (g == w).all(), (i, cws[i][1], len(cws[i][0]), int(np.argmax(g != w)))
