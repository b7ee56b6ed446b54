"""CPU tests: the oracle (oracle/oracle.c) against the committed golden fixtures.

tests/golden/*.npz were produced by tests/golden/make_golden.py from the UNMODIFIED reference
(oracle/_ref/libsrslte_ref.so compiled from /root/reference) and from the reference's own known-answer test
material (turbodecoder_test.h:70-125, crc_test.h:36-39).  These tests run anywhere gcc runs."""
import hashlib
import os

import numpy as np
import pytest

from oracle.bindings import CRC8, CRC16, CRC24A, CRC24B
from util import UL_GRANTS, all_K, lanes8, lanes16, ul_qprime

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], np.uint64)[0]


def test_kat_crc_words(port):
    """crc_test.h:36-39: CRC words of the srand(1) 5001-bit vector (bit-wise srslte_crc_checksum, crc_test.c:101)"""
    k = np.load(os.path.join(G, "kat.npz"))
    bits, words = k["crc_bits"], k["crc_words"]
    for (poly, order), want in zip(((CRC24A, 24), (CRC24B, 24), (CRC16, 16), (CRC8, 8)), words):
        assert port.crc_bits(poly, order, bits) == int(want), hex(poly)


def test_kat_turbo_encoder_vector(port):
    """turbodecoder_test.h:70-125: known_data -> known_data_encoded pins the QPP parameters of K=504, the constituent
    encoders and the tail-bit order the decoder's input extraction relies on"""
    k = np.load(os.path.join(G, "kat.npz"))
    got, want = port.tcod_encode(k["known_data"]), k["known_data_encoded"]
    # 1523 of the 1524 symbols agree.  Position 1512 (the first tail bit) of the reference's fixture differs from the
    # output of the reference's OWN srslte_tcod_encode (checked in test_oracle_vs_ref.py::test_known_vector_vs_reference_encoder);
    # turbodecoder_test only ever feeds the fixture to the decoder (turbodecoder_test.c:240), it never compares it.
    assert np.nonzero(got != want)[0].tolist() == [1512]


def test_kat_decodes_known_vector(port):
    """noise-free decode of the reference's known code word returns the known data (all decoder families that take K=504)"""
    k = np.load(os.path.join(G, "kat.npz"))
    cw = k["known_data_encoded"].astype(np.int16) * 200 - 100
    for dec, fnsb in ((0, True), (1, True), (3, True)):
        h = port.tdec_new(dec, fnsb)
        rc, out = port.tdec_run_all(h, cw, 2, 504)
        port.tdec_del(h)
        assert rc == 0 and (np.unpackbits(out) == k["known_data"]).all(), dec


def test_tables_against_reference_digests(port):
    t = np.load(os.path.join(G, "tables.npz"))
    Ks = t["K"].tolist()
    assert Ks == all_K() == [port.cbsize(i) for i in range(188)]
    qi = ri = 0
    for ci, K in enumerate(Ks):
        for lanes in (1, 8, 16, 32):
            if lanes > 1 and K % lanes:
                assert t["qpp"][qi] == 0
            else:
                f, r = port.qpp(K, lanes)
                assert digest(np.concatenate([f, r])) == t["qpp"][qi], (K, lanes)
            qi += 1
        for rv in range(4):
            e = (np.arange(3 * K + 12) + 1).astype(np.int16)
            for lanes in (0, lanes16(K)):
                out = np.zeros(18600, np.int16)
                port.rm_rx16(e, out, K, rv, lanes)
                assert digest(out) == t["rm"][ri], (K, rv, lanes)
                ri += 1
            e8 = ((np.arange(3 * K + 12) * 7 + 3) % 251 - 125).astype(np.int8)
            out8 = np.zeros(18600 * 2, np.int8)
            port.rm_rx8(e8, out8, K, rv, lanes8(K))
            assert digest(out8) == t["rm"][ri], (K, rv, "int8")
            ri += 1


def test_decoder_against_reference_outputs(port):
    t = np.load(os.path.join(G, "tdec.npz"))
    for idx, (K, bits, amp, nit) in enumerate(t["cases"].tolist()):
        llr = np.ascontiguousarray(t["in%d" % idx])
        h = port.tdec_new(0, False)
        assert port.tdec_new_cb(h, K) == 0
        for it in range(nit):
            out = port.tdec_iteration(h, llr, K)
            assert (out == t["bytes%d" % idx][it]).all(), (K, bits, amp, it)
            assert digest(port.tdec_get_llr(h, 2 if it % 2 == 0 else 0, K)) == t["llr%d" % idx][it], (K, bits, amp, it, "llr")
        port.tdec_del(h)


def test_transport_blocks_against_reference_outputs(port):
    t = np.load(os.path.join(G, "tb.npz"))
    for idx, c in enumerate(t["cases"].tolist()):
        tbs, Qm, Gb, bits, amp, sigma, mi = int(c[0]), int(c[1]), int(c[2]), int(c[3]), c[4], c[5], int(c[6])
        sb = port.softbuffer_new()
        for j in range(int(t["nrv"][idx])):
            rv, rc_ref = t["rv%d_%d" % (idx, j)].tolist()
            llr = np.ascontiguousarray(t["llr%d_%d" % (idx, j)])
            rc, d, nit, avg, crc = port.decode_tb(sb, tbs, Qm, rv, llr, mi)
            n = tbs // 8 + 6
            assert rc == rc_ref, (idx, j)
            assert (d[:n] == t["out%d_%d" % (idx, j)]).all(), (idx, j)
            assert abs(avg - float(t["avg%d_%d" % (idx, j)][0])) < 1e-6
            assert (crc == t["crc%d_%d" % (idx, j)]).all()
            if rc == 0:
                assert (d[:tbs // 8] == t["data%d" % idx]).all()
        port.softbuffer_del(sb)


def test_soft_demodulation_against_reference_outputs(port):
    """srslte_demod_soft_demodulate_{s,b}, srslte_sequence_LTE_pr and srslte_scrambling_{s,sb}_offset of the
    unmodified reference on fixed inputs (tests/golden/demod.npz)"""
    g = np.load(os.path.join(G, "demod.npz"))
    n_checked = 0
    for key in g.files:
        if not key.startswith("llr_"):
            continue
        _, mod, bits, n, a = key.split("_")
        got = port.demod(int(mod), g["sym_%s_%s" % (n, a)], np.int16 if bits == "16" else np.int8)
        assert (got == g[key]).all(), key
        n_checked += 1
    assert n_checked == 5 * 2 * 9 * 5
    for i, (c_init, L) in enumerate(g["seqs"]):
        cb = port.sequence_bytes(int(c_init), int(L))
        assert (cb[:int(L) // 8] == g["seq%d" % i]).all()
        for bits in (16, 8):
            assert (port.descramble(cb, g["din%d_%d" % (i, bits)]) == g["dout%d_%d" % (i, bits)]).all()


def test_ulsch_pre_steps_against_reference_outputs(port):
    """what the unmodified srslte_ulsch_decode leaves in g_bits / q_bits for fixed input LLRs
    (tests/golden/ulsch.npz): ACK zeroing, RI-skipping de-interleaver with its g_bits[0] side effect, CQI offset; and the
    transport block decoded from the restated g_bits."""
    u = np.load(os.path.join(G, "ulsch.npz"))
    for i, grant in enumerate(UL_GRANTS[:10]):
        tbs, Qm, L_prb, nof_symb = grant[:4]
        llr = u["llr%d" % i]
        qa, qr, qc = ul_qprime(port, grant)
        assert [qa, qr, qc] == u["qprime%d" % i].tolist()
        rc, g, ack, ri, q2 = port.ulsch_deinterleave(llr, Qm, nof_symb, qa, qr, g_fill=777)
        assert rc == 0
        assert digest(g[qc * Qm:]) == u["g_data%d" % i] and digest(q2) == u["q_after%d" % i], i
        front = g[:qc * Qm].copy()
        if grant[6] == 1 and len(front) > 32: # decode_cqi_short folds the copies onto the first 32 in place (uci.c:379-385)
            for k in range(1, len(front) // 32):
                front[:32] += front[32 * k:32 * k + 32]
            k = len(front) // 32
            front[:len(front) % 32] += front[32 * k:]
        assert (front == u["g_front%d" % i]).all(), i
        G_bits = (len(llr) // Qm - qr - qc) * Qm
        sb = port.softbuffer_new()
        rc_p, d, _, _, _ = port.decode_tb(sb, tbs, Qm, 0, g[qc * Qm:qc * Qm + G_bits].copy(), 8)
        port.softbuffer_del(sb)
        assert rc_p == int(u["rc%d" % i][0]) and (d[:tbs // 8] == u["data%d" % i]).all(), i
    # geometries the reference cannot index
    z = np.zeros(12 * 4 * 2, np.int16)
    assert port.ulsch_deinterleave(z, 2, 12, 17, 0)[0] == -1 and port.ulsch_deinterleave(z, 2, 12, 0, 17)[0] == -1
    assert port.ulsch_deinterleave(z[:-2], 2, 12, 0, 0)[0] == -1
