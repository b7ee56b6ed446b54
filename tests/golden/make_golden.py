#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference (oracle/_ref/libsrslte_ref.so, built by
oracle/build_ref.sh from /root/reference).  Run in the container that has /root/reference:

    python tests/golden/make_golden.py

Outputs (small, compressed):
  tests/golden/kat.npz       the reference's own known-answer material:
                             - turbodecoder_test.h:70-125 known_data[504] / known_data_encoded[1524]
                             - crc_test.h:36-39 expected CRC words of the srand(1) 5001-bit vector (+ the vector,
                               drawn with glibc rand() exactly as crc_test.c:86-93 does)
  tests/golden/tables.npz    digests of every QPP table and rate-dematch table of the reference (188 K x layouts x rv)
  tests/golden/tdec.npz      decided bytes after every half-iteration + LLR snapshots for a set of (K, width, amplitude)
  tests/golden/tb.npz        transport-block decodes through the real sch.c (incl. HARQ retransmissions)
  tests/golden/demod.npz     soft demodulation (all modulations, int16 / int8, lengths around the SIMD group sizes),
                             pseudo-random sequences and descrambled outputs
  tests/golden/ulsch.npz     PUSCH: input LLRs of srslte_ulsch_decode for the grants of tests/util.py UL_GRANTS[:10] (made by
                             srslte_ulsch_encode + AWGN) and what the call leaves behind: g_bits (digest of the UL-SCH part,
                             the CQI part in full), q_bits (digest), return code, decoded bytes
"""
import ctypes
import hashlib
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.bindings import CRC8, CRC16, CRC24A, CRC24B, Port, Ref, aligned_zeros  # noqa: E402
from util import UL_GRANTS, all_K, bpsk_awgn_llr, lanes8, lanes16, random_llr, ul_params, ul_qprime  # noqa: E402

REF_SRC = os.environ.get("SRSLTE_REFERENCE", "/root/reference")


def c_array(text, name):
    m = re.search(name + r"\[[^\]]*\]\s*=\s*\{(.*?)\};", text, re.S)
    return np.array([int(x) for x in m.group(1).replace("\n", " ").split(",") if x.strip()], np.uint8)


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], np.uint64)[0]


def main():
    R, P = Ref(), Port()
    # ---- known-answer material of the reference's own tests
    h = open(os.path.join(REF_SRC, "lib/src/phy/fec/test/turbodecoder_test.h")).read()
    known_data, known_enc = c_array(h, "known_data"), c_array(h, "known_data_encoded")
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    crc_bits = np.array([libc.rand() % 2 for _ in range(5001)], np.uint8)
    np.savez_compressed(os.path.join(HERE, "kat.npz"), known_data=known_data, known_data_encoded=known_enc, crc_bits=crc_bits,
                        crc_words=np.array([0x1C5C97, 0x36D1F0, 0x7FF4, 0xF0], np.uint32))
    # ---- tables
    Ks = all_K()
    qpp, rm = [], []
    for ci, K in enumerate(Ks):
        for lanes in (1, 8, 16, 32):
            if lanes > 1 and K % lanes:
                qpp.append(0)
                continue
            f, r = R.qpp(K, lanes)
            qpp.append(digest(np.concatenate([f, r])))
        for rv in range(4):
            e = (np.arange(3 * K + 12) + 1).astype(np.int16)
            out = np.zeros(18600, np.int16)
            R.rm_rx16(e, out, ci, rv, enable_sb=False)
            rm.append(digest(out))
            out = np.zeros(18600, np.int16)
            R.rm_rx16(e, out, ci, rv, enable_sb=True)
            rm.append(digest(out))
            e8 = ((np.arange(3 * K + 12) * 7 + 3) % 251 - 125).astype(np.int8)
            out8 = np.zeros(18600 * 2, np.int8)
            R.rm_rx8(e8, out8, ci, rv)
            rm.append(digest(out8))
    np.savez_compressed(os.path.join(HERE, "tables.npz"), K=np.array(Ks, np.uint32), qpp=np.array(qpp, np.uint64), rm=np.array(rm, np.uint64))
    # ---- decoder
    rng = np.random.default_rng(20201001)
    cases = [(6144, 16, 300, 4), (6144, 16, 30000, 4), (5824, 16, 2000, 6), (816, 16, 5000, 4), (512, 16, 20000, 4), (408, 16, 300, 3),
             (40, 16, 100, 4), (400, 16, 20000, 4), (6144, 8, 60, 4), (6144, 8, 127, 4), (2112, 8, 100, 3), (816, 8, 127, 4), (1008, 8, 127, 3),
             (400, 8, 127, 3), (512, 8, 60, 3)]
    out = {"cases": np.array(cases, np.int32)}
    for idx, (K, bits, amp, nit) in enumerate(cases):
        dt = np.int16 if bits == 16 else np.int8
        N = lanes16(K) if bits == 16 else lanes8(K)
        n = 3 * (K + 32) + 12 if N else 3 * K + 12
        llr = aligned_zeros(18600 * (2 if bits == 8 else 1), dt)
        llr[:n] = random_llr(rng, n, amp, dt)
        hd = R.tdec_new(0, False)
        assert R.tdec_new_cb(hd, K) == 0
        by, dg = [], []
        for it in range(nit):
            by.append(R.tdec_iteration(hd, llr, K, patched=True))
            dg.append(digest(R.tdec_get_llr(hd, 2 if it % 2 == 0 else 0, K)))
        R.tdec_del(hd)
        out["in%d" % idx] = llr[:n].copy()
        out["bytes%d" % idx] = np.stack(by)
        out["llr%d" % idx] = np.array(dg, np.uint64)
    np.savez_compressed(os.path.join(HERE, "tdec.npz"), **out)
    # ---- transport blocks through the real sch.c
    tbc = [(15264, 4, 20000, 16, 200, 0.55, 8, (0,)), (15264, 4, 20000, 16, 200, 0.75, 8, (0, 2)), (6120, 2, 14400, 16, 50, 0.9, 8, (0,)),
           (2216, 2, 7000, 8, 30, 0.6, 6, (0,)), (31704, 6, 40000, 16, 100, 0.5, 8, (0,)), (31704, 6, 40000, 8, 20, 0.62, 8, (0, 1))]
    out = {"cases": np.array([c[:7] for c in tbc], np.float64), "nrv": np.array([len(c[7]) for c in tbc], np.int32)}
    for idx, (tbs, Qm, G, bits, amp, sigma, mi, rvs) in enumerate(tbc):
        dt = np.int16 if bits == 16 else np.int8
        s = R.sch_new(bits == 8, mi, 100)
        R.sch_reset_rx(s, tbs)
        data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        out["data%d" % idx] = data
        for j, rv in enumerate(rvs):
            e = P.encode_tb(tbs, Qm, rv, G, data)  # (checked equal to the reference's encoder in tests/test_oracle_vs_ref.py)
            llr = bpsk_awgn_llr(rng, e, amp, sigma, dt)
            rc, d, avg, crc = R.sch_decode(s, tbs, Qm, rv, llr)
            out["llr%d_%d" % (idx, j)] = llr
            out["rv%d_%d" % (idx, j)] = np.array([rv, rc], np.int32)
            out["out%d_%d" % (idx, j)] = d[:tbs // 8 + 6].copy()
            out["avg%d_%d" % (idx, j)] = np.array([avg], np.float32)
            out["crc%d_%d" % (idx, j)] = crc.copy()
        R.sch_del(s)
    np.savez_compressed(os.path.join(HERE, "tb.npz"), **out)
    for f in ("kat", "tables", "tdec", "tb"):
        print(f, os.path.getsize(os.path.join(HERE, f + ".npz")))


def make_demod():
    """tests/golden/demod.npz: the reference's soft demodulator / sequence generator / descrambler on fixed inputs"""
    R = Ref()
    rng = np.random.default_rng(2026)
    out = {}
    lens = (1, 7, 8, 9, 16, 17, 31, 100, 203)
    amps = (0.3, 1.0, 3.0, 50.0, 400.0)
    for n in lens:
        for a, amp in enumerate(amps):
            out["sym_%d_%d" % (n, a)] = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * amp).astype(np.complex64)
    for mod in range(5):
        for bits, dt in ((16, np.int16), (8, np.int8)):
            for n in lens:
                for a in range(len(amps)):
                    out["llr_%d_%d_%d_%d" % (mod, bits, n, a)] = R.demod(mod, out["sym_%d_%d" % (n, a)], dt)
    seqs = ((1, 104), (12345, 9000), (0x7fffffff, 11520), ((61 << 14) + (3 << 9) + 150, 2880))
    out["seqs"] = np.array(seqs, np.int64)
    for i, (c_init, L) in enumerate(seqs):
        out["seq%d" % i] = R.sequence_bytes(c_init, L)[:L // 8]
        for bits, dt in ((16, np.int16), (8, np.int8)):
            d = rng.integers(np.iinfo(dt).min, np.iinfo(dt).max + 1, L).astype(dt)
            out["din%d_%d" % (i, bits)] = d
            out["dout%d_%d" % (i, bits)] = R.descramble(c_init, d)
    np.savez_compressed(os.path.join(HERE, "demod.npz"), **out)
    print("demod", os.path.getsize(os.path.join(HERE, "demod.npz")))


def make_ulsch():
    """tests/golden/ulsch.npz: the reference's srslte_ulsch_encode -> AWGN -> srslte_ulsch_decode on fixed inputs"""
    R, P = Ref(), Port()
    rng = np.random.default_rng(36212)
    out = {}
    for i, grant in enumerate(UL_GRANTS[:10]):
        tbs, Qm, L_prb, nof_symb = grant[:4]
        params = ul_params(grant)
        s = R.sch_new(False, 8, 100)
        data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        q_tx = R.ulsch_encode(s, params, int(rng.integers(0, 1 << 20)), data)
        llr = bpsk_awgn_llr(rng, q_tx, 100, 0.35, np.int16)
        c_seq = rng.integers(0, 2, len(llr), dtype=np.uint8)
        R.sch_reset_rx(s, tbs)
        rc, d, g, q_after, uci = R.ulsch_decode(s, params, llr, c_seq, g_fill=777)
        qa, qr, qc = ul_qprime(P, grant) # the reference does not report them; the fixture's g_bits only match if they are right
        out["llr%d" % i] = llr
        out["qprime%d" % i] = np.array([qa, qr, qc], np.uint32)
        out["g_front%d" % i] = g[:qc * Qm]
        out["g_data%d" % i] = digest(g[qc * Qm:])
        out["q_after%d" % i] = digest(q_after)
        out["rc%d" % i] = np.array([rc], np.int32)
        out["data%d" % i] = d[:tbs // 8]
        R.sch_del(s)
    np.savez_compressed(os.path.join(HERE, "ulsch.npz"), **out)
    print("ulsch", os.path.getsize(os.path.join(HERE, "ulsch.npz")))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "demod":
        make_demod()
    elif len(sys.argv) > 1 and sys.argv[1] == "ulsch":
        make_ulsch()
    else:
        main()
        make_demod()
        make_ulsch()
