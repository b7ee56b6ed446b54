"""srsran_b200 -- Python host mirror of the B200 LTE turbo-decode engine's C ABI.

The product is ``libsrslte_fec_b200.so`` (hand-written sm_100a CUDA behind the srsLTE FEC API, see
include/srslte_b200/*.h).  This module only binds it with ctypes so tests and bench.py can drive it; it contains
no arithmetic and has no CPU fallback: importing works anywhere, but creating a ``Context`` (or calling any
srslte_* symbol) without the built library or without a GPU raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SRSLTE_B200_LIB", os.path.join(HERE, "libsrslte_fec_b200.so"))

IN_DEVICE = 1
OUT_DEVICE = 2
UCI_DEFERRED = 8  # ulsch_deinterleave with OUT_DEVICE: the UCI LLR arrays are filled by the next wait() on the context
SEQ_DEVICE = 4
MAX_CODEBLOCKS = 32

TDEC_AUTO, TDEC_GENERIC, TDEC_SSE, TDEC_SSE_WINDOW, TDEC_NEON_WINDOW, TDEC_AVX_WINDOW, TDEC_SSE8_WINDOW, TDEC_AVX8_WINDOW = range(8)

CRC24A = 0x1864CFB
CRC24B = 0x1800063
CRC16 = 0x11021
CRC8 = 0x19B


class B200Error(RuntimeError):
    pass


class CbBatch(C.Structure):  # srslte_b200_cb_batch_t
    _fields_ = [("K", C.c_uint32), ("nof_cb", C.c_uint32), ("nof_iterations", C.c_uint32), ("llr_bits", C.c_uint32),
                ("llr_stride", C.c_uint32), ("input_sb", C.c_uint32), ("dec_type", C.c_uint32)]


class Tb(C.Structure):  # srslte_b200_tb_t
    _fields_ = [("e_bits", C.c_void_p), ("nof_e_bits", C.c_uint32), ("tbs", C.c_uint32), ("Qm", C.c_uint32), ("rv", C.c_uint32),
                ("softbuffer", C.c_void_p), ("data", C.c_void_p),
                ("ret", C.c_int32), ("avg_iterations", C.c_float), ("cb_crc", C.c_uint8 * MAX_CODEBLOCKS),
                ("cb_noi", C.c_uint8 * MAX_CODEBLOCKS), ("nof_cb", C.c_uint32)]


class Enc(C.Structure):  # srslte_b200_enc_t
    _fields_ = [("data", C.c_void_p), ("tbs", C.c_uint32), ("Qm", C.c_uint32), ("rv", C.c_uint32), ("nof_e_bits", C.c_uint32),
                ("e_bits", C.c_void_p), ("ret", C.c_int32)]


class Demod(C.Structure):  # srslte_b200_demod_t
    _fields_ = [("symbols", C.c_void_p), ("nof_symbols", C.c_uint32), ("mod", C.c_uint32), ("scramble_bytes", C.c_void_p), ("e_bits", C.c_void_p),
                ("csi", C.c_void_p)]


class Ulsch(C.Structure):  # srslte_b200_ulsch_t
    _fields_ = [("q_bits", C.c_void_p), ("Qm", C.c_uint32), ("H_prime_total", C.c_uint32), ("N_pusch_symbs", C.c_uint32),
                ("Q_prime_ack", C.c_uint32), ("Q_prime_ri", C.c_uint32), ("Q_prime_cqi", C.c_uint32), ("g_bits", C.c_void_p),
                ("ack_llr", C.c_void_p), ("ri_llr", C.c_void_p), ("cqi_llr", C.c_void_p)]


MOD_BITS = [1, 2, 4, 6, 8]  # bits per symbol of srslte_mod_t 0..4 (BPSK, QPSK, 16QAM, 64QAM, 256QAM)


class CbSegm(C.Structure):  # srslte_cbsegm_t
    _fields_ = [(n, C.c_uint32) for n in ("F", "C", "K1", "K2", "K1_idx", "K2_idx", "C1", "C2", "tbs")]


class Crc(C.Structure):  # srslte_crc_t
    _fields_ = [("table", C.c_uint64 * 256), ("polynom", C.c_int), ("order", C.c_int), ("crcinit", C.c_uint64),
                ("crcmask", C.c_uint64), ("crchighbit", C.c_uint64), ("srslte_crc_out", C.c_uint32)]


class Tdec(C.Structure):  # srslte_tdec_t
    _fields_ = [("max_long_cb", C.c_uint32), ("b200_engine", C.c_void_p), ("force_not_sb", C.c_bool), ("dec_type", C.c_int),
                ("current_long_cb", C.c_uint32), ("current_cbidx", C.c_int), ("n_iter", C.c_int)]




_lib = None


def lib():
    """Load libsrslte_fec_b200.so; fails loudly if it has not been built (python -m srsran_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error("libsrslte_fec_b200.so is missing: build it with `python -m srsran_b200.build` "
                            "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.srslte_b200_last_error.restype = C.c_char_p
        L.srslte_b200_host_alloc.restype = C.c_void_p
        L.srslte_b200_host_alloc.argtypes = [C.c_uint64]
        L.srslte_b200_host_alloc_wc.restype = C.c_void_p
        L.srslte_b200_host_alloc_wc.argtypes = [C.c_uint64]
        L.srslte_b200_host_free.argtypes = [C.c_void_p]
        L.srslte_b200_device_alloc.restype = C.c_void_p
        L.srslte_b200_device_alloc.argtypes = [C.c_uint64]
        L.srslte_b200_device_free.argtypes = [C.c_void_p]
        L.srslte_b200_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.srslte_b200_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.srslte_b200_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.srslte_b200_ctx_destroy.argtypes = [C.c_void_p]
        for f in ("srslte_b200_tdec_batch", "srslte_b200_tdec_batch_submit"):
            getattr(L, f).argtypes = [C.c_void_p, C.POINTER(CbBatch), C.c_void_p, C.c_void_p, C.c_uint32]
        for f in ("srslte_b200_decode_tbs", "srslte_b200_decode_tbs_submit"):
            getattr(L, f).argtypes = [C.c_void_p, C.POINTER(Tb), C.c_uint32, C.c_int, C.c_uint32, C.c_uint32]
        L.srslte_b200_wait.argtypes = [C.c_void_p]
        L.srslte_b200_demod_descramble.argtypes = [C.c_void_p, C.POINTER(Demod), C.c_uint32, C.c_int, C.c_uint32]
        L.srslte_b200_sequence_bytes.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
        L.srslte_b200_sequence_bytes.restype = None
        L.srslte_b200_encode_tbs.argtypes = [C.c_void_p, C.POINTER(Enc), C.c_uint32, C.c_uint32]
        L.srslte_b200_ulsch_deinterleave.argtypes = [C.c_void_p, C.POINTER(Ulsch), C.c_uint32, C.c_uint32]
        L.srslte_b200_softbuffer_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_uint32]
        L.srslte_b200_softbuffer_reset.argtypes = [C.c_void_p]
        L.srslte_b200_softbuffer_free.argtypes = [C.c_void_p]
        for f in ("srslte_b200_last_gpu_ms", "srslte_b200_last_map_ms"):
            getattr(L, f).restype = C.c_float
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("srslte_b200_last_launches", "srslte_b200_last_map_launches"):
            getattr(L, f).restype = C.c_uint32
            getattr(L, f).argtypes = [C.c_void_p]
        L.srslte_b200_timer_start.argtypes = [C.c_void_p]
        L.srslte_b200_timer_stop_ms.argtypes = [C.c_void_p]
        L.srslte_b200_timer_stop_ms.restype = C.c_float
        L.srslte_b200_alu_probe.argtypes = [C.c_void_p, C.c_int]
        L.srslte_b200_alu_probe.restype = C.c_double
        L.srslte_b200_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        for f in ("srslte_b200_last_replayed", "srslte_b200_last_half_iterations"):
            getattr(L, f).restype = C.c_uint32
            getattr(L, f).argtypes = [C.c_void_p]
        L.srslte_crc_checksum_byte.restype = C.c_uint32
        L.srslte_crc_checksum.restype = C.c_uint32
        L.srslte_crc_attach_byte.restype = C.c_uint32
        L.srslte_tdec_autoimp_get_subblocks.restype = C.c_uint32
        L.srslte_tdec_autoimp_get_subblocks_8bit.restype = C.c_uint32
        L.srslte_cbsegm_cbsize_isvalid.restype = C.c_bool
        _lib = L
    return _lib


def sequence_bytes(c_init, length):
    """TS 36.211 7.2 pseudo-random sequence packed like srslte_sequence_t::c_bytes (host helper of the library)"""
    out = np.zeros((length + 7) // 8, np.uint8)
    lib().srslte_b200_sequence_bytes(c_init, length, out.ctypes.data_as(C.c_void_p))
    return out


def make_demods(n):
    return (Demod * n)()


def make_encs(n):
    return (Enc * n)()


def make_ulschs(n):
    return (Ulsch * n)()


def _err(what, rc):
    raise B200Error("%s failed (%d): %s" % (what, rc, lib().srslte_b200_last_error().decode()))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class PinnedArray:
    """numpy view over page-locked host memory (srslte_b200_host_alloc)"""

    def __init__(self, shape, dtype, write_combined=False):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = (lib().srslte_b200_host_alloc_wc if write_combined else lib().srslte_b200_host_alloc)(max(self.nbytes, 16))
        if not self.ptr:
            raise B200Error("pinned host allocation failed")
        buf = (C.c_uint8 * max(self.nbytes, 16)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().srslte_b200_host_free(self.ptr)
            self.ptr = None


class Context:
    """One engine = one GPU + one stream (srslte_b200_ctx_t)."""

    def __init__(self, device=-1):
        self.h = C.c_void_p()
        rc = lib().srslte_b200_ctx_create(C.byref(self.h), device)
        if rc:
            _err("srslte_b200_ctx_create", rc)
        self._keep = None

    def close(self):
        if self.h:
            lib().srslte_b200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- device memory helpers
    def device_alloc(self, nbytes):
        p = lib().srslte_b200_device_alloc(nbytes)
        if not p:
            raise B200Error("device allocation of %d bytes failed" % nbytes)
        return p

    def device_free(self, p):
        lib().srslte_b200_device_free(p)

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        rc = lib().srslte_b200_memcpy_h2d(self.h, dptr, _ptr(arr), arr.nbytes)
        if rc:
            _err("h2d", rc)

    def d2h(self, arr, dptr):
        rc = lib().srslte_b200_memcpy_d2h(self.h, _ptr(arr), dptr, arr.nbytes)
        if rc:
            _err("d2h", rc)

    # ---- soft demodulation + descrambling (srslte_demod_soft_demodulate_{s,b} + srslte_scrambling_{s,sb}_offset)
    def demod_descramble(self, codewords, dtype):
        """codewords: list of (complex64 symbols, mod, packed sequence bytes or None[, float32 csi or None]), host arrays.
        Returns the list of int16 / int8 LLR arrays (nof_symbols * Qm each)."""
        n = len(codewords)
        arr = (Demod * n)()
        keep, outs = [], []
        for i, cw in enumerate(codewords):
            sym, mod, scr = cw[:3]
            csi = np.ascontiguousarray(cw[3], np.float32) if len(cw) > 3 and cw[3] is not None else None
            sym = np.ascontiguousarray(sym, np.complex64)
            out = np.zeros(len(sym) * MOD_BITS[mod], dtype)
            keep.append((sym, scr, csi))
            outs.append(out)
            arr[i].symbols, arr[i].nof_symbols, arr[i].mod = sym.ctypes.data, len(sym), mod
            arr[i].scramble_bytes = scr.ctypes.data if scr is not None else None
            arr[i].csi = csi.ctypes.data if csi is not None else None
            arr[i].e_bits = out.ctypes.data
        rc = lib().srslte_b200_demod_descramble(self.h, arr, n, int(dtype == np.int8), 0)
        if rc:
            _err("srslte_b200_demod_descramble", rc)
        return outs

    # ---- PUSCH pre-steps (the data movement of srslte_ulsch_decode before decode_tb)
    def ulsch_deinterleave(self, blocks, uci=True):
        """blocks: list of (q_bits int16[H' * Qm], Qm, N_pusch_symbs, Q'_ack, Q'_ri, Q'_cqi), host arrays.  Returns a list of
        (g_bits int16[(H' - Q'_ri) * Qm], ack_llr, ri_llr, cqi_llr)."""
        n = len(blocks)
        arr = (Ulsch * n)()
        keep, outs = [], []
        for i, (q, Qm, nsym, qa, qr, qc) in enumerate(blocks):
            q = np.ascontiguousarray(q, np.int16)
            H = len(q) // Qm
            o = (np.zeros((H - qr) * Qm, np.int16), np.zeros(qa * Qm, np.int16), np.zeros(qr * Qm, np.int16), np.zeros(qc * Qm, np.int16))
            keep.append(q)
            outs.append(o)
            a = arr[i]
            a.q_bits, a.Qm, a.H_prime_total, a.N_pusch_symbs, a.Q_prime_ack, a.Q_prime_ri, a.Q_prime_cqi = q.ctypes.data, Qm, H, nsym, qa, qr, qc
            a.g_bits = o[0].ctypes.data
            if uci:
                a.ack_llr, a.ri_llr, a.cqi_llr = o[1].ctypes.data, o[2].ctypes.data, o[3].ctypes.data
        rc = lib().srslte_b200_ulsch_deinterleave(self.h, arr, n, 0)
        if rc:
            _err("srslte_b200_ulsch_deinterleave", rc)
        return outs

    def ulsch_deinterleave_raw(self, arr, flags):
        """arr: ctypes array of Ulsch with caller-managed (host or device) pointers"""
        rc = lib().srslte_b200_ulsch_deinterleave(self.h, arr, len(arr), flags)
        if rc:
            _err("srslte_b200_ulsch_deinterleave", rc)

    # ---- transmit mirror (srslte_dlsch_encode2 semantics, stateless)
    def encode_tbs(self, blocks):
        """blocks: list of (payload bytes uint8[tbs/8], tbs, Qm, rv, G).  Returns (list of packed e-bit arrays, list of ret)."""
        n = len(blocks)
        arr = (Enc * n)()
        keep, outs = [], []
        for i, (data, tbs, Qm, rv, G) in enumerate(blocks):
            data = np.ascontiguousarray(data, np.uint8)
            out = np.zeros((G + 31) // 32 * 4, np.uint8)
            keep.append(data)
            outs.append(out)
            arr[i].data, arr[i].tbs, arr[i].Qm, arr[i].rv, arr[i].nof_e_bits, arr[i].e_bits = data.ctypes.data, tbs, Qm, rv, G, out.ctypes.data
        rc = lib().srslte_b200_encode_tbs(self.h, arr, n, 0)
        if rc:
            _err("srslte_b200_encode_tbs", rc)
        return [o[:(blocks[i][4] + 7) // 8] for i, o in enumerate(outs)], [arr[i].ret for i in range(n)]

    def encode_tbs_raw(self, arr, flags):
        rc = lib().srslte_b200_encode_tbs(self.h, arr, len(arr), flags)
        if rc:
            _err("srslte_b200_encode_tbs", rc)

    def demod_descramble_raw(self, arr, is8, flags):
        """arr: ctypes array of Demod with caller-managed (host or device) pointers"""
        rc = lib().srslte_b200_demod_descramble(self.h, arr, len(arr), int(is8), flags)
        if rc:
            _err("srslte_b200_demod_descramble", rc)

    # ---- batch of code blocks (srslte_tdec_run_all semantics)
    def tdec_batch(self, llr, K, nof_iterations, input_sb=False, dec_type=TDEC_AUTO, out=None):
        """llr: (nof_cb, stride) int16/int8 host array.  Returns (nof_cb, K/8) uint8."""
        assert llr.ndim == 2 and llr.flags.c_contiguous and llr.dtype in (np.int16, np.int8)
        ncb, stride = llr.shape
        cfg = CbBatch(K, ncb, nof_iterations, 16 if llr.dtype == np.int16 else 8, stride, int(input_sb), dec_type)
        if out is None:
            out = np.zeros((ncb, K // 8), np.uint8)
        rc = lib().srslte_b200_tdec_batch(self.h, C.byref(cfg), _ptr(llr), _ptr(out), 0)
        if rc:
            _err("srslte_b200_tdec_batch", rc)
        return out

    def tdec_batch_device(self, d_llr, d_out, K, ncb, stride, llr_bits, nof_iterations, input_sb=False, dec_type=TDEC_AUTO, submit_only=False):
        cfg = CbBatch(K, ncb, nof_iterations, llr_bits, stride, int(input_sb), dec_type)
        f = lib().srslte_b200_tdec_batch_submit if submit_only else lib().srslte_b200_tdec_batch
        rc = f(self.h, C.byref(cfg), d_llr, d_out, IN_DEVICE | OUT_DEVICE)
        if rc:
            _err("srslte_b200_tdec_batch", rc)

    def tdec_batch_submit(self, llr_ptr, out_ptr, K, ncb, stride, llr_bits, nof_iterations, input_sb=False, dec_type=TDEC_AUTO, flags=0):
        cfg = CbBatch(K, ncb, nof_iterations, llr_bits, stride, int(input_sb), dec_type)
        rc = lib().srslte_b200_tdec_batch_submit(self.h, C.byref(cfg), llr_ptr, out_ptr, flags)
        if rc:
            _err("srslte_b200_tdec_batch_submit", rc)

    def wait(self):
        rc = lib().srslte_b200_wait(self.h)
        if rc:
            _err("srslte_b200_wait", rc)

    # ---- batch of transport blocks (decode_tb semantics)
    def decode_tbs(self, tbs, llr_is_8bit, max_iterations, flags=0, submit_only=False):
        """tbs: ctypes array of Tb (inputs filled in); results are written into it."""
        f = lib().srslte_b200_decode_tbs_submit if submit_only else lib().srslte_b200_decode_tbs
        rc = f(self.h, tbs, len(tbs), int(llr_is_8bit), max_iterations, flags)
        if rc:
            _err("srslte_b200_decode_tbs", rc)

    def set_tb_hints(self, hints):
        """difficulty hints (float per transport block) for the next decode_tbs on this context: grouping only, never results"""
        h = np.ascontiguousarray(hints, np.float32)
        rc = lib().srslte_b200_set_tb_hints(self.h, h.ctypes.data_as(C.POINTER(C.c_float)), len(h))
        if rc:
            _err("srslte_b200_set_tb_hints", rc)

    def softbuffer_create(self, max_cb=MAX_CODEBLOCKS):
        sb = C.c_void_p()
        rc = lib().srslte_b200_softbuffer_create(self.h, C.byref(sb), max_cb)
        if rc:
            _err("srslte_b200_softbuffer_create", rc)
        return sb

    def softbuffer_reset(self, sb):
        lib().srslte_b200_softbuffer_reset(sb)

    def softbuffer_free(self, sb):
        lib().srslte_b200_softbuffer_free(sb)

    def set_option(self, name, value):
        rc = lib().srslte_b200_set_option(self.h, name.encode(), int(value))
        if rc:
            _err("srslte_b200_set_option(%s)" % name, rc)

    def debug_read_plane(self, cb, plane, n):
        """int16 LLR plane of a code block of the last completed batch (test hook, see batch.h)"""
        out = np.zeros(n, np.int16)
        f = lib().srslte_b200_debug_read_plane
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
        rc = f(self.h, cb, plane, _ptr(out), n)
        if rc:
            _err("srslte_b200_debug_read_plane", rc)
        return out

    def last_replayed(self):
        return lib().srslte_b200_last_replayed(self.h)

    def last_half_iterations(self):
        return lib().srslte_b200_last_half_iterations(self.h)

    # ---- measurement hooks
    def timer_start(self):
        rc = lib().srslte_b200_timer_start(self.h)
        if rc:
            _err("srslte_b200_timer_start", rc)

    def timer_stop_ms(self):
        return lib().srslte_b200_timer_stop_ms(self.h)

    def alu_probe(self, mode=0):
        """packed int16x2 operations per second (see include/srslte_b200/batch.h)"""
        return lib().srslte_b200_alu_probe(self.h, mode)

    def last_gpu_ms(self):
        return lib().srslte_b200_last_gpu_ms(self.h)

    def last_map_ms(self):
        return lib().srslte_b200_last_map_ms(self.h)

    def last_launches(self):
        return lib().srslte_b200_last_launches(self.h)

    def last_map_launches(self):
        return lib().srslte_b200_last_map_launches(self.h)


def make_tbs(n):
    return (Tb * n)()


def qpp_table(K, lanes=1):
    f = np.zeros(K, np.uint16)
    r = np.zeros(K, np.uint16)
    rc = lib().srslte_b200_qpp_table(C.c_uint32(K), C.c_uint32(lanes), _ptr(f), _ptr(r))
    if rc:
        raise B200Error("srslte_b200_qpp_table(%d, %d) failed" % (K, lanes))
    return f, r


def rm_table(K, rv, lanes=0):
    t = np.zeros(3 * K + 12, np.uint16)
    rc = lib().srslte_b200_rm_table(C.c_uint32(K), C.c_uint32(rv), C.c_uint32(lanes), _ptr(t))
    if rc:
        raise B200Error("srslte_b200_rm_table(%d, %d, %d) failed" % (K, rv, lanes))
    return t


# ------------------------------------------------------------------ drop-in srslte_* symbols (host pointers)
def cbsegm(tbs):
    s = CbSegm()
    r = lib().srslte_cbsegm(C.byref(s), C.c_uint32(tbs))
    return r, {n: getattr(s, n) for n, _ in CbSegm._fields_}


def crc_checksum_byte(poly, order, data):
    c = Crc()
    assert lib().srslte_crc_init(C.byref(c), C.c_uint32(poly), C.c_int(order)) == 0
    d = np.ascontiguousarray(data, np.uint8)
    return lib().srslte_crc_checksum_byte(C.byref(c), _ptr(d), C.c_int(8 * len(d)))


def crc_checksum_bits(poly, order, bits):
    c = Crc()
    assert lib().srslte_crc_init(C.byref(c), C.c_uint32(poly), C.c_int(order)) == 0
    d = np.ascontiguousarray(bits, np.uint8)
    return lib().srslte_crc_checksum(C.byref(c), _ptr(d), C.c_int(len(d)))


def rm_turbo_rx_lut(e, out, cb_idx, rv, enable_input_tdec=True):
    """srslte_rm_turbo_rx_lut_ / _8bit on host arrays (accumulates into `out` in place)."""
    assert e.dtype == out.dtype and e.flags.c_contiguous and out.flags.c_contiguous
    if e.dtype == np.int16:
        return lib().srslte_rm_turbo_rx_lut_(_ptr(e), _ptr(out), C.c_uint32(len(e)), C.c_uint32(cb_idx), C.c_uint32(rv), C.c_bool(enable_input_tdec))
    return lib().srslte_rm_turbo_rx_lut_8bit(_ptr(e), _ptr(out), C.c_uint32(len(e)), C.c_uint32(cb_idx), C.c_uint32(rv))


class TurboDecoder:
    """srslte_tdec_t through the drop-in symbols."""

    def __init__(self, max_long_cb=6144, dec_type=TDEC_AUTO, force_not_sb=False):
        self.h = Tdec()
        rc = lib().srslte_tdec_init_manual(C.byref(self.h), C.c_uint32(max_long_cb), C.c_int(dec_type))
        if rc:
            raise B200Error("srslte_tdec_init_manual failed: %s" % lib().srslte_b200_last_error().decode())
        if force_not_sb:
            lib().srslte_tdec_force_not_sb(C.byref(self.h))

    def new_cb(self, K):
        return lib().srslte_tdec_new_cb(C.byref(self.h), C.c_uint32(K))

    def iteration(self, llr, K):
        out = np.zeros(K // 8, np.uint8)
        if llr.dtype == np.int16:
            lib().srslte_tdec_iteration(C.byref(self.h), _ptr(llr), _ptr(out))
        else:
            lib().srslte_tdec_iteration_8bit(C.byref(self.h), _ptr(llr), _ptr(out))
        return out

    def run_all(self, llr, nof_iterations, K):
        out = np.zeros(K // 8, np.uint8)
        f = lib().srslte_tdec_run_all if llr.dtype == np.int16 else lib().srslte_tdec_run_all_8bit
        rc = f(C.byref(self.h), _ptr(llr), _ptr(out), C.c_uint32(nof_iterations), C.c_uint32(K))
        return rc, out

    def n_iter(self):
        return lib().srslte_tdec_get_nof_iterations(C.byref(self.h))

    def free(self):
        lib().srslte_tdec_free(C.byref(self.h))


class SoftbufferRx:
    """srslte_softbuffer_rx_t through the drop-in symbols (device-resident HARQ buffer)"""

    def __init__(self, nof_prb=100):
        self.q = SoftbufferRx_t()
        rc = lib().srslte_softbuffer_rx_init(C.byref(self.q), C.c_uint32(nof_prb))
        if rc:
            raise B200Error("srslte_softbuffer_rx_init failed: %s" % lib().srslte_b200_last_error().decode())

    def reset_tbs(self, tbs):
        lib().srslte_softbuffer_rx_reset_tbs(C.byref(self.q), C.c_uint32(tbs))

    def cb_crc(self, n):
        return [bool(self.q.cb_crc[i]) for i in range(n)]

    def free(self):
        lib().srslte_softbuffer_rx_free(C.byref(self.q))


SoftbufferRx_t = SoftbufferRx  # placeholder, replaced below


class _SoftbufferRxStruct(C.Structure):  # srslte_softbuffer_rx_t (the reference's fields, then the device handle)
    _fields_ = [("max_cb", C.c_uint32), ("buffer_f", C.POINTER(C.POINTER(C.c_int16))), ("data", C.POINTER(C.POINTER(C.c_uint8))),
                ("cb_crc", C.POINTER(C.c_bool)), ("tb_crc", C.c_bool), ("b200_softbuffer", C.c_void_p), ("b200_host_dirty", C.c_bool)]


SoftbufferRx_t = _SoftbufferRxStruct


def decode_tb(softbuffer, tbs, Qm, rv, e_bits, max_iterations):
    """srslte_b200_decode_tb: decode_tb of sch.c:503-570 as one call on host buffers.
    Returns (ret, data bytes, avg_iterations)."""
    seg = CbSegm()
    if lib().srslte_cbsegm(C.byref(seg), C.c_uint32(tbs)):
        return -1, None, 0.0
    data = np.zeros(tbs // 8 + 8 + 768, np.uint8)
    avg = C.c_float(0)
    f = lib().srslte_b200_decode_tb
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_bool, C.c_uint32, C.c_void_p, C.c_void_p]
    rc = f(C.byref(softbuffer.q), C.byref(seg), Qm, rv, len(e_bits), _ptr(e_bits), e_bits.dtype == np.int8, max_iterations, _ptr(data), C.byref(avg))
    return rc, data, avg.value
