"""CPU tests: the oracle restatement against the UNMODIFIED reference compiled here (oracle/_ref).  Skipped where the
reference library does not exist (it is built only where /root/reference is mounted; the golden tests cover the rest)."""
import numpy as np
import pytest

from oracle.bindings import (CRC8, CRC16, CRC24A, CRC24B, TDEC_AUTO, TDEC_AVX8_WINDOW, TDEC_AVX_WINDOW, TDEC_GENERIC, TDEC_SSE8_WINDOW,
                             TDEC_SSE_WINDOW, MOD_BITS, aligned_zeros)
from util import all_K, bpsk_awgn_llr, random_llr, UL_GRANTS, ul_params, ul_qprime


def test_segmentation_and_dispatch(port, ref):
    for tbs in list(range(16, 100000, 8 * 53)) + [75376, 97896, 6120, 6128, 40, 75000]:
        assert port.cbsegm(tbs) == ref.cbsegm(tbs), tbs
    for i, K in enumerate(all_K()):
        assert port.cbsize(i) == ref.cbsize(i) == K
        assert port.cbindex(K) == ref.cbindex(K) == i
        assert port.subblocks16(K) == ref.subblocks16(K) and port.subblocks8(K) == ref.subblocks8(K)


def test_qpp_all_sizes(port, ref):
    for K in all_K():
        for lanes in (1, 8, 16, 32):
            if lanes > 1 and K % lanes:
                continue
            f, r = port.qpp(K, lanes)
            f2, r2 = ref.qpp(K, lanes)
            assert (f == f2).all() and (r == r2).all(), (K, lanes)


def test_crc(port, ref):
    rng = np.random.default_rng(1)
    for n in (1, 3, 100, 768, 9422):
        d = rng.integers(0, 256, n, dtype=np.uint8)
        for poly, o in ((CRC24A, 24), (CRC24B, 24), (CRC16, 16), (CRC8, 8)):
            assert port.crc_bytes(poly, o, d) == ref.crc_bytes(poly, o, d)
    for n in (5001, 40, 41, 47):
        b = rng.integers(0, 2, n, dtype=np.uint8)
        for poly, o in ((CRC24A, 24), (CRC24B, 24), (CRC16, 16), (CRC8, 8)):
            assert port.crc_bits(poly, o, b) == ref.crc_bits(poly, o, b)


def test_rate_dematching_tables_all_sizes(port, ref):
    rng = np.random.default_rng(2)
    for ci, K in enumerate(all_K()):
        for rv in range(4):
            E = 3 * K + 12
            e = (np.arange(E) + 1).astype(np.int16)
            for sb in (False, True):
                lanes = port.subblocks16(K) if sb else 0
                out = np.zeros(18600, np.int16)
                ref.rm_rx16(e, out, ci, rv, enable_sb=sb)
                want = np.zeros(18600, np.int16)
                want[port.rm_table(K, rv, lanes)] = e
                assert (out == want).all(), (K, rv, sb)
    for K, E in ((40, 500), (6144, 30000), (1024, 1000), (5824, 6918)):
        ci = port.cbindex(K)
        for rv in range(4):
            e = random_llr(rng, E, 32767, np.int16)
            a = random_llr(rng, 18600, 32767, np.int16)
            b = a.copy()
            ref.rm_rx16(e, a, ci, rv, True)
            port.rm_rx16(e, b, K, rv, port.subblocks16(K))
            assert (a == b).all()
            e = random_llr(rng, E, 127, np.int8)
            a = random_llr(rng, 18600 * 2, 127, np.int8)
            b = a.copy()
            ref.rm_rx8(e, a, ci, rv)
            port.rm_rx8(e, b, K, rv, port.subblocks8(K))
            assert (a == b).all()


def test_rm_lut_equals_procedural_float(ref):
    """what rm_turbo_test.c:171-190 checks (there only for cb_idx 0, rv 0): the LUT de-matcher in standard layout equals
    the procedural float de-matcher exactly"""
    rng = np.random.default_rng(3)
    for ci in (0, 45, 77, 100, 187):
        K = ref.cbsize(ci)
        for rv in range(4):
            for E in (1920, 8192, 3 * K + 12):
                e = random_llr(rng, E, 100, np.int16)
                out = np.zeros(18600, np.int16)
                ref.rm_rx16(e, out, ci, rv, enable_sb=False)
                f = ref.rm_rx_float(e.astype(np.float32), 3 * K + 12, rv)
                f[f == 10000] = 0  # SRSLTE_RX_NULL marks positions that received nothing
                assert (out[:3 * K + 12] == f.astype(np.int16)).all(), (ci, rv, E)


def _tdec(port, ref, K, dtype, amp, nit, dec=TDEC_AUTO, fnsb=False, seed=0):
    rng = np.random.default_rng(seed + K)
    bits = 16 if dtype == np.int16 else 8
    if dec == TDEC_AUTO:
        N = port.subblocks16(K) if bits == 16 else port.subblocks8(K)
        sb_in = N > 0 and not fnsb
    else:
        sb_in = False
    n = 3 * (K + 32) + 12 if sb_in else 3 * K + 12
    a = aligned_zeros(18600 * (2 if bits == 8 else 1), dtype)
    a[:n] = random_llr(rng, n, amp, dtype)
    b = a.copy()
    hr, hp = ref.tdec_new(dec, fnsb), port.tdec_new(dec, fnsb)
    assert ref.tdec_new_cb(hr, K) == 0 and port.tdec_new_cb(hp, K) == 0
    for it in range(nit):
        o1, o2 = ref.tdec_iteration(hr, a, K), port.tdec_iteration(hp, b, K)
        for w in ((2,) if it % 2 == 0 else (0, 2)):
            assert (ref.tdec_get_llr(hr, w, K) == port.tdec_get_llr(hp, w, K)).all(), (K, bits, amp, it, w)
        assert (o1 == o2).all(), (K, bits, amp, it)
    ref.tdec_del(hr)
    port.tdec_del(hp)


@pytest.mark.parametrize("case", [(6144, 16, 300, 4), (6144, 16, 30000, 6), (816, 16, 5000, 6), (512, 16, 20000, 6), (408, 16, 300, 5), (800, 16, 32767, 4),
                                  (5824, 16, 2000, 10), (6144, 8, 60, 6), (6144, 8, 127, 6), (2112, 8, 100, 5), (816, 8, 127, 6), (2048, 8, 90, 6),
                                  (1008, 8, 127, 4), (5824, 8, 50, 10), (40, 16, 100, 6), (400, 16, 300, 6), (400, 16, 20000, 6), (200, 16, 32767, 5),
                                  (104, 16, 1000, 10), (40, 8, 100, 4), (400, 8, 127, 6), (408, 8, 127, 4), (512, 8, 60, 6), (800, 8, 127, 5)])
def test_decoder_llr_exact(port, ref, case):
    K, bits, amp, nit = case
    _tdec(port, ref, K, np.int16 if bits == 16 else np.int8, amp, nit)


@pytest.mark.parametrize("bits,amp", [(16, 700), (16, 32767), (8, 40), (8, 127)])
def test_decoder_llr_exact_all_sizes(port, ref, bits, amp):
    """the oracle's decoders pinned LLR-for-LLR against the reference on EVERY one of the 188 LTE sizes (AUTO dispatch:
    generic, 8-, 16- and 32-lane windowed decoders), moderate and full-scale amplitudes, 3 half-iterations each"""
    for K in all_K():
        _tdec(port, ref, K, np.int16 if bits == 16 else np.int8, amp, 3)


@pytest.mark.parametrize("K,dec,fnsb", [(6144, TDEC_AVX_WINDOW, True), (504, TDEC_SSE_WINDOW, True), (504, TDEC_GENERIC, True), (6144, TDEC_GENERIC, False),
                                        (1024, TDEC_SSE_WINDOW, False), (6144, TDEC_AVX8_WINDOW, True), (1024, TDEC_SSE8_WINDOW, True), (504, TDEC_AUTO, True),
                                        (6144, TDEC_AUTO, True), (40, TDEC_AUTO, True)])
def test_decoder_manual_and_force_not_sb(port, ref, K, dec, fnsb):
    _tdec(port, ref, K, np.int16, 400, 4, dec, fnsb)


def test_encoder_chain(port, ref):
    rng = np.random.default_rng(7)
    for K in (40, 504, 6144, 1024):
        b = rng.integers(0, 2, K, dtype=np.uint8)
        assert (port.tcod_encode(b) == ref.tcod_encode(b)).all()
    s = ref.sch_new(False, 8, 100)
    for tbs, Qm, G in ((75376, 6, 90000), (97896, 8, 115200), (6120, 2, 14400), (15264, 4, 20000)):
        for rv in range(4):
            d = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
            ref.sch_encode(s, tbs, Qm, G, 0, d)
            assert (ref.sch_encode(s, tbs, Qm, G, rv, d) == port.encode_tb(tbs, Qm, rv, G, d)).all(), (tbs, rv)
    ref.sch_del(s)


@pytest.mark.parametrize("tbs,Qm,G,is8,amp,sigma,mi,rvs", [
    (75376, 6, 90000, False, 100, 0.0, 8, (0,)), (75376, 6, 90000, False, 100, 0.46, 8, (0,)), (75376, 6, 90000, False, 100, 0.95, 8, (0, 2, 3, 1)),
    (97896, 8, 115200, True, 20, 0.42, 8, (0,)), (97896, 8, 115200, True, 20, 0.8, 8, (0, 2)), (6120, 2, 14400, False, 50, 0.9, 8, (0,)),
    (296, 2, 1200, True, 30, 0.5, 8, (0,)), (15264, 4, 20000, False, 200, 0.7, 10, (0, 1))])
def test_transport_block_through_real_sch(port, ref, tbs, Qm, G, is8, amp, sigma, mi, rvs):
    """decode_tb / decode_tb_cb restated vs the reference's real srslte_dlsch_decode2 (sch.c), incl. HARQ combining"""
    rng = np.random.default_rng(tbs + G)
    s, sb = ref.sch_new(is8, mi, 100), port.softbuffer_new()
    ref.sch_reset_rx(s, tbs)
    data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
    for rv in rvs:
        ref.sch_encode(s, tbs, Qm, G, 0, data)
        e = ref.sch_encode(s, tbs, Qm, G, rv, data)
        llr = bpsk_awgn_llr(rng, e, amp, sigma, np.int8 if is8 else np.int16)
        rc1, d1, avg1, crc1 = ref.sch_decode(s, tbs, Qm, rv, llr)
        rc2, d2, nit, avg2, crc2 = port.decode_tb(sb, tbs, Qm, rv, llr, mi)
        n = tbs // 8 + 3
        assert rc1 == rc2 and (d1[:n] == d2[:n]).all() and abs(avg1 - avg2) < 1e-6 and (crc1 == crc2).all(), (tbs, rv)
        if rc1 == 0:
            assert (d1[:tbs // 8] == data).all()
            break
    ref.sch_del(s)
    port.softbuffer_del(sb)


def test_known_vector_vs_reference_encoder(ref):
    """the reference's known_data_encoded fixture vs the reference's own encoder: one tail bit (index 1512) differs"""
    import os
    k = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat.npz"))
    got = ref.tcod_encode(k["known_data"])
    assert np.nonzero(got != k["known_data_encoded"])[0].tolist() == [1512]


# ----------------------------------------------------------------------------------------- SURVEY 8f row 1
@pytest.mark.parametrize("mod", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("dtype", [np.int16, np.int8])
def test_soft_demodulation(port, ref, mod, dtype):
    """srslte_demod_soft_demodulate_{s,b}: the restatement reproduces the reference's SSE bodies (round to nearest,
    saturating packs, integer thresholds) AND its scalar tails (truncation, wrap, float thresholds) -- lengths around
    the SIMD group sizes, amplitudes up to saturation"""
    rng = np.random.default_rng(100 + mod)
    for n in (1, 3, 4, 7, 8, 9, 15, 16, 17, 31, 32, 100, 1001, 15000):
        for amp in (0.3, 1.0, 3.0, 50.0, 400.0):
            sym = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * amp).astype(np.complex64)
            assert (port.demod(mod, sym, dtype) == ref.demod(mod, sym, dtype)).all(), (n, amp)


def test_sequence_and_descrambling(port, ref):
    """srslte_sequence_LTE_pr (c_bytes) and srslte_scrambling_{s,sb}_offset"""
    rng = np.random.default_rng(5)
    for c_init, L in ((1, 104), (12345, 90000), (0x7fffffff, 115200), (777, 40), ((61 << 14) + (3 << 9) + 150, 28800)):
        cb = port.sequence_bytes(c_init, L)
        assert (cb[:L // 8] == ref.sequence_bytes(c_init, L)[:L // 8]).all()
        for dtype in (np.int16, np.int8):
            d = rng.integers(np.iinfo(dtype).min, np.iinfo(dtype).max + 1, L).astype(dtype)
            assert (port.descramble(cb, d) == ref.descramble(c_init, d)).all()


# ----------------------------------------------------------------------------------------- SURVEY 8f rank 2
@pytest.mark.parametrize("grant", UL_GRANTS)
def test_ulsch_pre_steps(port, ref, grant):
    """The real srslte_ulsch_encode -> AWGN -> srslte_ulsch_decode (sch.c:1105-1330) against the restatement of its data
    movement: g_bits after the call (incl. the RI-clobbered g_bits[0] and the untouched tail), q_bits after the call (ACK
    positions zeroed) -- which also pins the Q' formulas and the ACK / RI positions -- and the transport block decoded
    from the restated g_bits."""
    tbs, Qm, L_prb, nof_symb = grant[:4]
    rng = np.random.default_rng(tbs + 7 * nof_symb + grant[4])
    params = ul_params(grant)
    s = ref.sch_new(False, 8, 100)
    data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
    q_tx = ref.ulsch_encode(s, params, int(rng.integers(0, 1 << 20)), data)
    nb_q = L_prb * 12 * nof_symb * Qm
    assert len(q_tx) == nb_q
    c_seq = rng.integers(0, 2, nb_q, dtype=np.uint8)
    llr = bpsk_awgn_llr(rng, q_tx, 100, 0.35, np.int16)
    ref.sch_reset_rx(s, tbs)
    rc_ref, d_ref, g_ref, q_after, out = ref.ulsch_decode(s, params, llr, c_seq, g_fill=777)
    qa, qr, qc = ul_qprime(port, grant)
    assert (qa > 0) == (grant[4] > 0) and (qr > 0) == (grant[5] > 0) and (qc > 0) == (grant[6] > 0)
    rc, g, ack, ri, q2 = port.ulsch_deinterleave(llr, Qm, nof_symb, qa, qr, g_fill=777)
    assert rc == 0
    assert (q2 == q_after).all()
    Qc = qc * Qm
    assert (g[Qc:] == g_ref[Qc:]).all()
    if qr:
        assert (g[-qr * Qm:] == 777).all()
    # the CQI LLRs at the front: the reference's short-CQI decoder folds them onto the first 32 in place (decode_cqi_short,
    # uci.c:379-385) after the de-interleaver wrote them; its long-CQI decoder only reads them
    front = g[:Qc].copy()
    if grant[6] == 1 and Qc > 32:
        for i in range(1, Qc // 32):
            front[:32] += front[32 * i:32 * i + 32]
        i = max(Qc // 32, 1)
        front[:Qc % 32] += front[32 * i:32 * i + Qc % 32]
    assert (front == g_ref[:Qc]).all()
    if qr and not qc:
        assert g[0] == llr[max(port.ulsch_uci_position(True, i, Qm, nb_q // Qm, nof_symb) for i in range(qr)) + Qm - 1]
    for i in range(qa):
        p = port.ulsch_uci_position(False, i, Qm, nb_q // Qm, nof_symb)
        assert (ack[i * Qm:(i + 1) * Qm] == llr[p:p + Qm]).all() and (q2[p:p + Qm] == 0).all()
    for i in range(qr):
        p = port.ulsch_uci_position(True, i, Qm, nb_q // Qm, nof_symb)
        assert (ri[i * Qm:(i + 1) * Qm] == llr[p:p + Qm]).all()
    # the transport block from the restated g_bits through the restated decode_tb
    G = (nb_q // Qm - qr - qc) * Qm
    sb = port.softbuffer_new()
    rc_p, d_p, _, _, _ = port.decode_tb(sb, tbs, Qm, 0, g[qc * Qm:qc * Qm + G].copy(), 8)
    port.softbuffer_del(sb)
    assert rc_p == rc_ref
    assert (d_p[:tbs // 8] == d_ref[:tbs // 8]).all()
    if grant != UL_GRANTS[8]: # that one loses a third of its resource elements to a 126x ACK offset and does not decode
        assert rc_ref == 0 and (d_p[:tbs // 8] == data).all()
    ref.sch_del(s)


@pytest.mark.parametrize("dtype", [np.int16, np.int8])
def test_csi_correction(port, ref, dtype):
    """SURVEY 8f row 1, third item: csi_correction (pdsch.c:628-741, static -- run through oracle/ref_csi_shim.c): every
    modulation, lengths around its SSE group sizes, full-range soft bits, csi values incl. equal maxima and a zero"""
    rng = np.random.default_rng(628 + (dtype == np.int8))
    info = np.iinfo(dtype)
    for mod in range(5):
        qm = MOD_BITS[mod]
        # (the reference's loop bounds `i < nof_bits - 3` / `- 11` are unsigned: a one-symbol QPSK / 64QAM codeword runs off its
        #  arrays there; real grants have hundreds of symbols, so n starts at 2)
        for n in (2, 3, 4, 5, 7, 8, 9, 16, 17, 31, 100, 1001, 14400):
            csi = rng.random(n).astype(np.float32) * (1.7 if n % 2 else 0.4) + 0.01
            if n > 4:
                csi[3] = csi.max()  # the maximum twice
                csi[1] = 0.0
            e = rng.integers(info.min, info.max + 1, n * qm).astype(dtype)
            want = ref.csi_correction(csi, e, mod)
            got = port.csi_correction(csi, e, mod)
            assert (got == want).all(), (mod, n, int(np.argmax(got != want)))


def test_latency_probes_leave_the_thread_affinity_alone(ref):
    """the per-TTI latency probes pin the calling thread to one core; threads created afterwards inherit the caller's mask,
    so a probe that does not put it back serialises every later multi-threaded measurement on one core"""
    import os
    before = os.sched_getaffinity(0)
    llr = aligned_zeros(2 * 1568, np.int16).reshape(2, 1568)   # row stride a multiple of 32 bytes
    llr[:] = np.random.default_rng(5).integers(-60, 60, llr.shape)
    lat = ref.latency_c1(llr, 512, 2, 2, 3)
    assert len(lat) == 3 and (lat > 0).all()
    assert os.sched_getaffinity(0) == before
    tb = np.random.default_rng(6).integers(-60, 60, (1, 1200)).astype(np.int16)
    lat = ref.latency_tb(tb, 1000, 2, 4, 3)
    assert len(lat) == 3 and os.sched_getaffinity(0) == before
