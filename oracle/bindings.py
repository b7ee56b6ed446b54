"""ctypes bindings for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Two libraries:
  * ``Port``  -> oracle/liboracle.so      (oracle.c, our scalar restatement; always buildable with gcc)
  * ``Ref``   -> oracle/_ref/libsrslte_ref.so (the UNMODIFIED reference compiled by oracle/build_ref.sh
                 from /root/reference; exists only where it was built -- it travels to the GPU box as a
                 prebuilt file)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsrslte_ref.so")
REF_FAST_SO = os.path.join(HERE, "_ref", "libsrslte_ref_fast.so")

CRC24A = 0x1864CFB
CRC24B = 0x1800063
CRC16 = 0x11021
CRC8 = 0x19B

# srslte_tdec_impl_type_t (turbodecoder_impl.h:28-38)
TDEC_AUTO, TDEC_GENERIC, TDEC_SSE, TDEC_SSE_WINDOW, TDEC_NEON_WINDOW, TDEC_AVX_WINDOW, TDEC_SSE8_WINDOW, TDEC_AVX8_WINDOW = range(8)
# srslte_mod_t (phy_common.h): bits per symbol
MOD_BPSK, MOD_QPSK, MOD_16QAM, MOD_64QAM, MOD_256QAM = range(5)
MOD_BITS = [1, 2, 4, 6, 8]


def build(ref=True):
    """Compile liboracle.so (always) and the reference library (if /root/reference exists)."""
    subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if ref and os.path.isdir(os.environ.get("SRSLTE_REFERENCE", "/root/reference")):
        subprocess.check_call([os.path.join(HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)
        if os.path.exists(os.path.join(os.path.dirname(HERE), "srsran_b200", "libsrslte_fec_b200.so")):
            subprocess.check_call([os.path.join(HERE, "build_opt_a.sh")], stdout=subprocess.DEVNULL)


def aligned_zeros(n, dtype, align=64):
    """The reference decodes in place from its input with aligned SIMD loads (turbodecoder_win.h:633):
    LLR buffers handed to Ref.tdec_* must be 32-byte aligned."""
    dtype = np.dtype(dtype)
    raw = np.zeros(n * dtype.itemsize + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n * dtype.itemsize].view(dtype)


def _p(a, t=None):
    return a.ctypes.data_as(C.c_void_p)


def _i16(x):
    return np.ascontiguousarray(x, dtype=np.int16)


def _i8(x):
    return np.ascontiguousarray(x, dtype=np.int8)


def _u8(x):
    return np.ascontiguousarray(x, dtype=np.uint8)


class Port:
    """oracle.c"""

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(ref=False)
        L = self.L = C.CDLL(PORT_SO)
        L.orc_tdec_new.restype = C.c_void_p
        L.orc_softbuffer_new.restype = C.c_void_p
        L.orc_crc_bytes.restype = C.c_uint32
        L.orc_crc_bits.restype = C.c_uint32
        L.orc_subblocks16.restype = C.c_uint32
        L.orc_subblocks8.restype = C.c_uint32

    # tables
    def cbsegm(self, tbs):
        out = np.zeros(9, np.uint32)
        r = self.L.orc_cbsegm(C.c_uint32(tbs), _p(out))
        return r, dict(zip(["F", "C", "K1", "K2", "K1_idx", "K2_idx", "C1", "C2", "tbs"], out.tolist()))

    def cbsize(self, idx):
        return self.L.orc_cbsize(C.c_uint32(idx))

    def cbindex(self, K):
        return self.L.orc_cbindex(C.c_uint32(K))

    def subblocks16(self, K):
        return self.L.orc_subblocks16(C.c_uint32(K))

    def subblocks8(self, K):
        return self.L.orc_subblocks8(C.c_uint32(K))

    def qpp(self, K, nsb):
        f = np.zeros(K, np.uint16)
        r = np.zeros(K, np.uint16)
        rc = self.L.orc_qpp(C.c_uint32(K), C.c_uint32(nsb), _p(f), _p(r))
        assert rc == 0
        return f, r

    def rm_table(self, K, rv, nsb):
        t = np.zeros(3 * K + 12, np.uint16)
        self.L.orc_rm_table(C.c_uint32(K), C.c_uint32(rv), C.c_uint32(nsb), _p(t))
        return t

    def rm_rx16(self, e, out, K, rv, nsb):
        e = _i16(e)
        assert out.dtype == np.int16 and out.flags.c_contiguous
        return self.L.orc_rm_rx16(_p(e), _p(out), C.c_uint32(len(e)), C.c_uint32(K), C.c_uint32(rv), C.c_uint32(nsb))

    def rm_rx8(self, e, out, K, rv, nsb):
        e = _i8(e)
        assert out.dtype == np.int8 and out.flags.c_contiguous
        return self.L.orc_rm_rx8(_p(e), _p(out), C.c_uint32(len(e)), C.c_uint32(K), C.c_uint32(rv), C.c_uint32(nsb))

    def crc_bytes(self, poly, order, data):
        d = _u8(data)
        return self.L.orc_crc_bytes(C.c_uint32(poly), C.c_int(order), _p(d), C.c_uint32(len(d)))

    def crc_bits(self, poly, order, bits):
        d = _u8(bits)
        return self.L.orc_crc_bits(C.c_uint32(poly), C.c_int(order), _p(d), C.c_uint32(len(d)))

    # transmit side (test-input synthesis)
    def tcod_encode(self, bits):
        b = _u8(bits)
        out = np.zeros(3 * len(b) + 12, np.uint8)
        rc = self.L.orc_tcod_encode(_p(b), _p(out), C.c_uint32(len(b)))
        assert rc == 0
        return out

    def encode_tb(self, tbs, Qm, rv, G, data):
        d = _u8(data)
        assert len(d) >= tbs // 8
        e = np.zeros(G, np.uint8)
        rc = self.L.orc_encode_tb(C.c_uint32(tbs), C.c_uint32(Qm), C.c_uint32(rv), C.c_uint32(G), _p(d), _p(e))
        assert rc == 0, rc
        return e

    def map_win(self, bits, N, x, apr, par, K):
        out = np.zeros(K + 3, np.int16)
        self.L.orc_map_win(C.c_int(bits), C.c_uint32(N), _p(_i16(x)), _p(_i16(apr)) if apr is not None else None, _p(_i16(par)), _p(out), C.c_uint32(K))
        return out[:K]

    # decoder object
    def tdec_new(self, dec_type=TDEC_AUTO, force_not_sb=False):
        return C.c_void_p(self.L.orc_tdec_new(C.c_int(dec_type), C.c_int(int(force_not_sb))))

    def tdec_del(self, h):
        self.L.orc_tdec_del(h)

    def tdec_new_cb(self, h, K):
        return self.L.orc_tdec_new_cb(h, C.c_uint32(K))

    def tdec_iteration(self, h, llr, K):
        """llr: np.int16 or np.int8 array; returns K/8 decided bytes"""
        assert llr.dtype in (np.int16, np.int8) and llr.flags.c_contiguous
        out = np.zeros(K // 8, np.uint8)
        self.L.orc_tdec_iteration(h, _p(llr), C.c_int(16 if llr.dtype == np.int16 else 8), _p(out))
        return out

    def tdec_run_all(self, h, llr, nof_iter, K):
        assert llr.dtype in (np.int16, np.int8) and llr.flags.c_contiguous
        out = np.zeros(K // 8, np.uint8)
        rc = self.L.orc_tdec_run_all(h, _p(llr), C.c_int(16 if llr.dtype == np.int16 else 8), _p(out), C.c_uint32(nof_iter), C.c_uint32(K))
        return rc, out

    def tdec_get_llr(self, h, which, n):
        d = np.zeros(n, np.int16)
        self.L.orc_tdec_get_llr(h, C.c_int(which), _p(d), C.c_uint32(n))
        return d

    def tdec_n_iter(self, h):
        return self.L.orc_tdec_n_iter(h)

    # TB level
    def softbuffer_new(self):
        return C.c_void_p(self.L.orc_softbuffer_new())

    def softbuffer_del(self, s):
        self.L.orc_softbuffer_del(s)

    def softbuffer_reset(self, s):
        self.L.orc_softbuffer_reset(s)

    def softbuffer_get(self, s, cb, n=18600):
        d = np.zeros(n, np.int16)
        self.L.orc_softbuffer_get(s, C.c_uint32(cb), _p(d), C.c_uint32(n))
        return d

    def decode_tb(self, s, tbs, Qm, rv, e_bits, max_iter, ncb_hint=32):
        assert e_bits.dtype in (np.int16, np.int8) and e_bits.flags.c_contiguous
        data = np.zeros(tbs // 8 + 8 + 768, np.uint8)
        nit = np.zeros(32, np.uint32)
        avg = C.c_float(0)
        rc = self.L.orc_decode_tb(s, C.c_uint32(tbs), C.c_uint32(Qm), C.c_uint32(rv), C.c_uint32(len(e_bits)), _p(e_bits),
                                  C.c_int(int(e_bits.dtype == np.int8)), C.c_uint32(max_iter), _p(data), _p(nit), C.byref(avg))
        crc = np.zeros(32, np.uint8)
        self.L.orc_softbuffer_get_crc(s, _p(crc), C.c_uint32(32))
        return rc, data, nit, avg.value, crc

    # ---- soft demodulation + descrambling
    def demod(self, mod, symbols, dtype):
        """symbols: complex64[n]; returns int16/int8 [n * Qm] (srslte_demod_soft_demodulate_{s,b})"""
        sym = np.ascontiguousarray(symbols, np.complex64).view(np.float32)
        n = len(sym) // 2
        out = np.zeros(n * MOD_BITS[mod], dtype)
        f = self.L.orc_demod_s if dtype == np.int16 else self.L.orc_demod_b
        assert f(C.c_int(mod), _p(sym), _p(out), C.c_int(n)) == 0
        return out

    def csi_correction(self, csi, e, mod):
        """csi_correction (pdsch.c:628-741) on a copy of e (int16 / int8 soft bits of one codeword)"""
        out = np.ascontiguousarray(e).copy()
        c = np.ascontiguousarray(csi, np.float32)
        self.L.orc_csi_correction(_p(c), _p(out), C.c_uint32(len(out)), C.c_int(mod), C.c_int(int(out.dtype == np.int8)))
        return out

    def sequence_bytes(self, c_init, length):
        out = np.zeros((length + 7) // 8, np.uint8)
        self.L.orc_sequence_bytes(C.c_uint32(c_init), C.c_uint32(length), _p(out))
        return out

    def descramble(self, c_bytes, data):
        d = data.copy()
        (self.L.orc_descramble_s if d.dtype == np.int16 else self.L.orc_descramble_b)(_p(c_bytes), _p(d), C.c_int(len(d)))
        return d

    # ---- PUSCH pre-steps
    def ulsch_qprime(self, K_segm, L_prb, nof_symb, nof_ack, ri_len, cqi_len, I_ack, I_ri, I_cqi):
        """(Q'_ack, Q'_ri, Q'_cqi) as srslte_ulsch_decode derives them for a grant that carries UL-SCH data (sch.c:1021-1180)"""
        L = self.L
        L.orc_beta_offset.restype = C.c_float
        L.orc_qprime_ri_ack.restype = C.c_uint32
        L.orc_qprime_cqi.restype = C.c_uint32
        u = C.c_uint32
        beta = lambda w, i: C.c_float(L.orc_beta_offset(C.c_int(w), u(i)))
        qa = L.orc_qprime_ri_ack(u(K_segm), u(L_prb), u(nof_symb), u(nof_ack), u(cqi_len), beta(0, I_ack)) if nof_ack else 0
        qr = L.orc_qprime_ri_ack(u(K_segm), u(L_prb), u(nof_symb), u(ri_len), u(cqi_len), beta(1, I_ri)) if ri_len else 0
        qc = L.orc_qprime_cqi(u(K_segm), u(L_prb), u(nof_symb), u(cqi_len), beta(2, I_cqi), u(qr)) if cqi_len else 0
        return qa, qr, qc

    def ulsch_uci_position(self, is_ri, idx, Qm, H_prime_total, N_pusch_symbs):
        self.L.orc_ulsch_uci_position.restype = C.c_int64
        return self.L.orc_ulsch_uci_position(C.c_int(int(is_ri)), C.c_uint32(idx), C.c_uint32(Qm), C.c_uint32(H_prime_total), C.c_uint32(N_pusch_symbs))

    def ulsch_deinterleave(self, q_bits, Qm, N_pusch_symbs, Q_prime_ack, Q_prime_ri, g_fill=0):
        """q_bits int16[H' * Qm].  Returns (rc, g_bits, ack_llr, ri_llr, q_bits after the ACK positions were zeroed)"""
        q = _i16(q_bits).copy()
        H = len(q) // Qm
        g = np.full(len(q), g_fill, np.int16)
        ack = np.zeros(max(Q_prime_ack * Qm, 1), np.int16)
        ri = np.zeros(max(Q_prime_ri * Qm, 1), np.int16)
        rc = self.L.orc_ulsch_deinterleave(_p(q), C.c_uint32(Qm), C.c_uint32(H), C.c_uint32(N_pusch_symbs), C.c_uint32(Q_prime_ack),
                                           C.c_uint32(Q_prime_ri), _p(g), _p(ack), _p(ri))
        return rc, g, ack[:Q_prime_ack * Qm], ri[:Q_prime_ri * Qm], q


class Ref:
    """The unmodified reference (srsLTE 20.10.1) through oracle/ref_shim.c"""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self, fast=False):
        """fast=True: the build with the reference's release flags (-Ofast -funroll-loops), for bench.py's CPU baseline only"""
        so = REF_SO
        if fast and os.path.exists(REF_FAST_SO):
            so = REF_FAST_SO
        if not os.path.exists(so):
            raise RuntimeError("reference library not built: run oracle/build_ref.sh where /root/reference exists")
        self.flags = "-Ofast -funroll-loops -mavx2 -mfma" if so == REF_FAST_SO else "-O3 -mavx2 -mfma"
        L = self.L = C.CDLL(so)
        L.ref_bench_tb_mixed.restype = C.c_double
        L.ref_tdec_new.restype = C.c_void_p
        L.ref_sch_new.restype = C.c_void_p
        L.ref_crc_byte.restype = C.c_uint32
        L.ref_crc_bits.restype = C.c_uint32
        L.ref_subblocks16.restype = C.c_uint32
        L.ref_subblocks8.restype = C.c_uint32
        L.ref_bench_c1.restype = C.c_double
        L.ref_bench_tb.restype = C.c_double
        L.ref_init()

    def cbsegm(self, tbs):
        out = np.zeros(9, np.uint32)
        r = self.L.ref_cbsegm(C.c_uint32(tbs), _p(out))
        return r, dict(zip(["F", "C", "K1", "K2", "K1_idx", "K2_idx", "C1", "C2", "tbs"], out.tolist()))

    def cbsize(self, idx):
        return self.L.ref_cbsize(C.c_uint32(idx))

    def cbindex(self, K):
        return self.L.ref_cbindex(C.c_uint32(K))

    def subblocks16(self, K):
        return self.L.ref_subblocks16(C.c_uint32(K))

    def subblocks8(self, K):
        return self.L.ref_subblocks8(C.c_uint32(K))

    def qpp(self, K, nsb):
        f = np.zeros(K, np.uint16)
        r = np.zeros(K, np.uint16)
        rc = self.L.ref_qpp(C.c_uint32(K), C.c_uint32(max(nsb, 1)), _p(f), _p(r))
        assert rc == 0
        return f, r

    def crc_bytes(self, poly, order, data):
        d = _u8(data).copy()
        return self.L.ref_crc_byte(C.c_uint32(poly), C.c_int(order), _p(d), C.c_int(8 * len(d)))

    def crc_bits(self, poly, order, bits):
        d = _u8(bits).copy()
        return self.L.ref_crc_bits(C.c_uint32(poly), C.c_int(order), _p(d), C.c_int(len(d)))

    def rm_rx16(self, e, out, cb_idx, rv, enable_sb=True):
        e = _i16(e).copy()
        assert out.dtype == np.int16 and out.flags.c_contiguous
        return self.L.ref_rm_rx_lut16(_p(e), _p(out), C.c_uint32(len(e)), C.c_uint32(cb_idx), C.c_uint32(rv), C.c_int(int(enable_sb)))

    def rm_rx8(self, e, out, cb_idx, rv):
        e = _i8(e).copy()
        assert out.dtype == np.int8 and out.flags.c_contiguous
        return self.L.ref_rm_rx_lut8(_p(e), _p(out), C.c_uint32(len(e)), C.c_uint32(cb_idx), C.c_uint32(rv))

    def rm_tx(self, bits, E, rv):
        b = _u8(bits).copy()
        out = np.zeros(E, np.uint8)
        rc = self.L.ref_rm_tx(_p(b), C.c_uint32(len(b)), _p(out), C.c_uint32(E), C.c_uint32(rv))
        assert rc == 0
        return out

    def rm_rx_float(self, e, out_len, rv):
        e = np.ascontiguousarray(e, np.float32).copy()
        out = np.zeros(out_len, np.float32)
        rc = self.L.ref_rm_rx_float(_p(e), C.c_uint32(len(e)), _p(out), C.c_uint32(out_len), C.c_uint32(rv))
        assert rc == 0
        return out

    def tcod_encode(self, bits):
        b = _u8(bits).copy()
        K = len(b)
        out = np.zeros(3 * K + 12, np.uint8)
        rc = self.L.ref_tcod_encode(_p(b), _p(out), C.c_uint32(K))
        assert rc == 0
        return out

    def tdec_new(self, dec_type=TDEC_AUTO, force_not_sb=False, max_k=6144):
        h = self.L.ref_tdec_new(C.c_uint32(max_k), C.c_int(dec_type), C.c_int(int(force_not_sb)))
        assert h
        return C.c_void_p(h)

    def tdec_del(self, h):
        self.L.ref_tdec_del(h)

    def tdec_new_cb(self, h, K):
        return self.L.ref_tdec_new_cb(h, C.c_uint32(K))

    def tdec_iteration(self, h, llr, K, patched=True):
        """NOTE: the reference decodes in place from `llr` (it is the HARQ soft buffer); pass the same array each call"""
        assert llr.dtype in (np.int16, np.int8) and llr.flags.c_contiguous and llr.ctypes.data % 32 == 0
        out = np.zeros(K // 8, np.uint8)
        if llr.dtype == np.int16:
            self.L.ref_tdec_iteration16(h, _p(llr), _p(out))
        else:
            self.L.ref_tdec_iteration8(h, _p(llr), _p(out), C.c_int(int(patched)))
        return out

    def tdec_run_all(self, h, llr, nof_iter, K):
        assert llr.dtype in (np.int16, np.int8) and llr.flags.c_contiguous and llr.ctypes.data % 32 == 0
        out = np.zeros(K // 8, np.uint8)
        if llr.dtype == np.int16:
            rc = self.L.ref_tdec_run_all16(h, _p(llr), _p(out), C.c_uint32(nof_iter), C.c_uint32(K))
        else:
            rc = self.L.ref_tdec_run_all8(h, _p(llr), _p(out), C.c_uint32(nof_iter), C.c_uint32(K))
        return rc, out

    def tdec_get_llr(self, h, which, n):
        d = np.zeros(n, np.int16)
        self.L.ref_tdec_get_llr(h, C.c_int(which), _p(d), C.c_uint32(n))
        return d

    def tdec_n_iter(self, h):
        return self.L.ref_tdec_n_iter(h)

    def tdec_current(self, h):
        o = np.zeros(3, np.int32)
        self.L.ref_tdec_current(h, _p(o))
        return o.tolist()

    # TB level through the real sch.c
    def sch_new(self, llr_is_8bit=False, max_iter=10, nof_prb=100):
        s = self.L.ref_sch_new(C.c_int(int(llr_is_8bit)), C.c_uint32(max_iter), C.c_uint32(nof_prb))
        assert s
        return C.c_void_p(s)

    def sch_del(self, s):
        self.L.ref_sch_del(s)

    def sch_reset_rx(self, s, tbs):
        self.L.ref_sch_reset_rx(s, C.c_uint32(tbs))

    def sch_encode(self, s, tbs, Qm, G, rv, data):
        d = np.zeros(tbs // 8 + 16, np.uint8)
        d[: tbs // 8] = _u8(data)[: tbs // 8]
        e = np.zeros((G + 7) // 8 + 64, np.uint8)
        rc = self.L.ref_sch_encode(s, C.c_uint32(tbs), C.c_uint32(Qm), C.c_uint32(G), C.c_uint32(rv), _p(d), _p(e))
        assert rc == 0, rc
        return np.unpackbits(e)[:G]

    def sch_decode(self, s, tbs, Qm, rv, llr):
        assert llr.dtype in (np.int16, np.int8) and llr.flags.c_contiguous
        data = np.zeros(tbs // 8 + 8 + 768, np.uint8)
        avg = C.c_float(0)
        crc = np.zeros(32, np.uint8)
        rc = self.L.ref_sch_decode(s, C.c_uint32(tbs), C.c_uint32(Qm), C.c_uint32(len(llr)), C.c_uint32(rv), _p(llr), _p(data),
                                   C.byref(avg), _p(crc), C.c_uint32(32))
        return rc, data, avg.value, crc

    def sch_get_softbuffer(self, s, cb, n=18600):
        d = np.zeros(n, np.int16)
        rc = self.L.ref_sch_get_softbuffer(s, C.c_uint32(cb), _p(d), C.c_uint32(n))
        assert rc == 0
        return d

    # ---- PUSCH through the real srslte_ulsch_encode / srslte_ulsch_decode (sch.c:1105-1330)
    @staticmethod
    def _ul_params(tbs, Qm, L_prb, nof_symb, rv, nof_ack, ri_len, cqi, I_ack, I_ri, I_cqi):
        return np.array([tbs, Qm, L_prb, nof_symb, rv, nof_ack, ri_len, cqi, I_ack, I_ri, I_cqi], np.uint32)

    def ulsch_encode(self, s, params, uci_seed, data):
        p = self._ul_params(*params)
        nb_q = int(p[2]) * 12 * int(p[3]) * int(p[1])
        d = np.zeros(int(p[0]) // 8 + 16, np.uint8)
        d[: int(p[0]) // 8] = _u8(data)[: int(p[0]) // 8]
        q = np.zeros(nb_q // 8 + 64, np.uint8)
        rc = self.L.ref_ulsch_encode(s, _p(p), C.c_uint32(uci_seed), _p(d), _p(q))
        assert rc >= 0, rc
        return np.unpackbits(q)[:nb_q]

    def ulsch_decode(self, s, params, q_llr, c_seq_bits, g_fill=0):
        """Returns (rc, data, g_bits, q_llr after the call, out[9] = ret, ack[0..3], ack.valid, ri, cqi crc, wideband cqi)"""
        p = self._ul_params(*params)
        q = aligned_zeros(len(q_llr) + 64, np.int16)
        q[:len(q_llr)] = q_llr
        g = aligned_zeros(len(q_llr) + 64, np.int16)
        g[:] = g_fill
        c = np.zeros(len(q_llr) + 64, np.uint8)
        c[:len(q_llr)] = c_seq_bits
        data = np.zeros(int(p[0]) // 8 + 8 + 768, np.uint8)
        out = np.zeros(16, np.int32)
        rc = self.L.ref_ulsch_decode(s, _p(p), _p(q), _p(c), _p(g), _p(data), _p(out))
        return rc, data, g[:len(q_llr)].copy(), q[:len(q_llr)].copy(), out

    # CPU baseline runners
    # ---- soft demodulation + descrambling; the SSE bodies use aligned loads/stores
    def demod(self, mod, symbols, dtype):
        n = len(symbols)
        sym = aligned_zeros(2 * n + 16, np.float32)
        sym[:2 * n] = np.ascontiguousarray(symbols, np.complex64).view(np.float32)
        out = aligned_zeros(n * MOD_BITS[mod] + 64, dtype)
        f = self.L.ref_demod_s if dtype == np.int16 else self.L.ref_demod_b
        assert f(C.c_int(mod), _p(sym), _p(out), C.c_int(n)) == 0
        return out[:n * MOD_BITS[mod]].copy()

    def csi_correction(self, csi, e, mod):
        """the reference's static csi_correction (pdsch.c:628-741) through oracle/ref_csi_shim.c, on a copy of e"""
        L = C.CDLL(os.path.join(HERE, "_ref", "libsrslte_ref_csi.so"))
        out = aligned_zeros(len(e) + 64, e.dtype)
        out[:len(e)] = e
        c = aligned_zeros(len(csi) + 16, np.float32)
        c[:len(csi)] = csi
        L.ref_csi_correction(_p(c), _p(out), C.c_uint32(len(e)), C.c_int(mod), C.c_int(int(e.dtype == np.int8)))
        return out[:len(e)].copy()

    def sequence_bytes(self, c_init, length):
        out = np.zeros((length + 7) // 8 + 16, np.uint8)
        assert self.L.ref_sequence_bytes(C.c_uint32(c_init), C.c_uint32(length), _p(out)) == 0
        return out[:(length + 7) // 8].copy()

    def descramble(self, c_init, data):
        d = aligned_zeros(len(data) + 64, data.dtype)
        d[:len(data)] = data
        f = self.L.ref_descramble_s if data.dtype == np.int16 else self.L.ref_descramble_b
        assert f(C.c_uint32(c_init), _p(d), C.c_int(len(data))) == 0
        return d[:len(data)].copy()

    def bench_c1(self, nthreads, llr, K, nof_iter, layout_sb=False, repeat=1):
        """llr: (ncb, stride) int16/int8.  Returns (seconds, out bytes (ncb, K/8))."""
        assert llr.ndim == 2 and llr.flags.c_contiguous
        ncb, stride = llr.shape
        out = np.zeros((ncb, K // 8), np.uint8)
        t = self.L.ref_bench_c1(C.c_int(nthreads), _p(llr), C.c_uint32(stride), C.c_uint32(ncb), C.c_uint32(K), C.c_uint32(nof_iter),
                                C.c_int(int(llr.dtype == np.int8)), C.c_int(int(layout_sb)), _p(out), C.c_uint32(repeat))
        return t, out

    def bench_tb_mixed(self, nthreads, llr_flat, off, tbs, Qm, G, max_iter, repeat=1):
        """transport blocks of different sizes: llr_flat holds all e-bits, off[i] the element offset of block i.
        Returns (seconds, data (ntb, stride), rc, avg_iter)."""
        ntb = len(tbs)
        off = np.ascontiguousarray(off, np.uint64)
        tbs, Qm, G = (np.ascontiguousarray(x, np.uint32) for x in (tbs, Qm, G))
        stride = int(tbs.max()) // 8 + 8 + 768
        out = np.zeros((ntb, stride), np.uint8)
        rc = np.zeros(ntb, np.int32)
        avg = np.zeros(ntb, np.float32)
        t = self.L.ref_bench_tb_mixed(C.c_int(nthreads), _p(llr_flat), _p(off), _p(tbs), _p(Qm), _p(G), C.c_uint32(ntb), C.c_uint32(max_iter),
                                      C.c_int(int(llr_flat.dtype == np.int8)), _p(out), C.c_uint32(stride), _p(rc), _p(avg), C.c_uint32(repeat))
        return t, out, rc, avg

    def bench_tb_symbols(self, nthreads, sym, mod, tbs, c_init, max_iter, is8, repeat=1):
        """sym: (ntb, nsym) complex64 equalised symbols; per block: soft demodulation + descrambling + srslte_dlsch_decode2.
        Returns (seconds, rc, avg_iter)."""
        ntb, nsym = sym.shape
        s = aligned_zeros(ntb * nsym * 2 + 16, np.float32)
        s[:ntb * nsym * 2] = np.ascontiguousarray(sym, np.complex64).view(np.float32).reshape(-1)
        stride = tbs // 8 + 8 + 768
        out = np.zeros((ntb, stride), np.uint8)
        rc = np.zeros(ntb, np.int32)
        avg = np.zeros(ntb, np.float32)
        self.L.ref_bench_tb_symbols.restype = C.c_double
        t = self.L.ref_bench_tb_symbols(C.c_int(nthreads), _p(s), C.c_uint32(ntb), C.c_uint32(nsym), C.c_uint32(mod), C.c_uint32(tbs), C.c_uint32(c_init),
                                        C.c_uint32(max_iter), C.c_int(int(is8)), _p(out), C.c_uint32(stride), _p(rc), _p(avg), C.c_uint32(repeat))
        return t, rc, avg

    def latency_tb(self, llr, tbs, Qm, max_iter, n_calls):
        """per-call wall time (us) of srslte_dlsch_decode2 for one transport block on one pinned core"""
        ntb, G = llr.shape
        lat = np.zeros(n_calls, np.float64)
        self.L.ref_latency_tb(_p(llr), C.c_uint32(ntb), C.c_uint32(tbs), C.c_uint32(Qm), C.c_uint32(G), C.c_uint32(max_iter),
                              C.c_int(int(llr.dtype == np.int8)), C.c_uint32(n_calls), _p(lat))
        return lat

    def latency_c1(self, llr, K, nof_iter, cbs_per_tti, n_calls):
        """per-call wall time (us) of cbs_per_tti x srslte_tdec_run_all on one pinned core"""
        ncb, stride = llr.shape
        lat = np.zeros(n_calls, np.float64)
        self.L.ref_latency_c1(_p(llr), C.c_uint32(stride), C.c_uint32(ncb), C.c_uint32(K), C.c_uint32(nof_iter), C.c_uint32(cbs_per_tti),
                              C.c_uint32(n_calls), _p(lat))
        return lat

    def bench_tb(self, nthreads, llr, tbs, Qm, rv, max_iter, repeat=1):
        """llr: (ntb, G) int16/int8.  Returns (seconds, data (ntb, stride), rc (ntb,), avg_iter (ntb,))."""
        assert llr.ndim == 2 and llr.flags.c_contiguous
        ntb, G = llr.shape
        stride = tbs // 8 + 8 + 768
        out = np.zeros((ntb, stride), np.uint8)
        rc = np.zeros(ntb, np.int32)
        avg = np.zeros(ntb, np.float32)
        t = self.L.ref_bench_tb(C.c_int(nthreads), _p(llr), C.c_uint32(ntb), C.c_uint32(tbs), C.c_uint32(Qm), C.c_uint32(G), C.c_uint32(rv),
                                C.c_uint32(max_iter), C.c_int(int(llr.dtype == np.int8)), _p(out), C.c_uint32(stride), _p(rc), _p(avg),
                                C.c_uint32(repeat))
        return t, out, rc, avg


OPT_A_SO = os.path.join(HERE, "_ref", "libsch_on_b200.so")


class OptA:
    """The reference's UNMODIFIED sch.c (decode_tb_cb's per-code-block loop) linked against the B200 library instead of the
    reference's decoder, rate de-matcher, CRC, segmentation and soft buffer (oracle/build_opt_a.sh; INTEGRATION.md option A)."""

    @staticmethod
    def available():
        return os.path.exists(OPT_A_SO)

    def __init__(self, llr_is_8bit=False, max_iter=8):
        L = self.L = C.CDLL(OPT_A_SO)
        L.opta_new.restype = C.c_void_p
        self.h = C.c_void_p(L.opta_new(C.c_int(int(llr_is_8bit)), C.c_uint32(max_iter)))
        if not self.h:
            raise RuntimeError("srslte_sch_init failed on the B200 library")

    def close(self):
        if self.h:
            self.L.opta_del(self.h)
            self.h = None

    def reset_rx(self, tbs):
        self.L.opta_reset_rx(self.h, C.c_uint32(tbs))

    def encode(self, tbs, Qm, G, rv, data):
        d = np.zeros(tbs // 8 + 16, np.uint8)  # (encode_tb appends the 3 CRC24A bytes to the caller's buffer, sch.c:216)
        d[:tbs // 8] = data
        e = np.zeros((G + 7) // 8 + 64, np.uint8)
        rc = self.L.opta_encode(self.h, C.c_uint32(tbs), C.c_uint32(Qm), C.c_uint32(G), C.c_uint32(rv), _p(d), _p(e))
        assert rc == 0
        return np.unpackbits(e)[:G]

    def decode(self, tbs, Qm, rv, llr):
        out = np.zeros(tbs // 8 + 8 + 768, np.uint8)
        avg = C.c_float(0)
        crc = np.zeros(32, np.uint8)
        llr = np.ascontiguousarray(llr)
        rc = self.L.opta_decode(self.h, C.c_uint32(tbs), C.c_uint32(Qm), C.c_uint32(len(llr)), C.c_uint32(rv), _p(llr), _p(out), C.byref(avg), _p(crc),
                                C.c_uint32(32))
        return rc, out, avg.value, crc
