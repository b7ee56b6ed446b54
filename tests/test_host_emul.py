"""CPU tests of the CUDA kernel's ALGORITHM: srsran_b200/csrc/map_core.cuh + arith.cuh are __host__ __device__ and are
compiled here with g++ (tests/host_emul/emul_map.cpp).  The emulation runs the very schedule the GPU runs -- chunked
passes, beta checkpoints + register-segment recompute, lane exchange, Fast16 range monitor -- for the T = N/2 "threads"
of one code block and is compared with the oracle's MAP call."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from util import random_llr

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "emul_map.cpp")
SO = os.path.join(HERE, "host_emul", "libemul_map.so")
CSRC = os.path.join(os.path.dirname(HERE), "srsran_b200", "csrc")


@pytest.fixture(scope="module")
def emul():
    deps = [SRC, os.path.join(CSRC, "map_core.cuh"), os.path.join(CSRC, "arith.cuh"), os.path.join(CSRC, "map_f16.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", SO, SRC])
    return C.CDLL(SO)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


CASES = [(16, 16, 16, 6144, 300, 0), (16, 16, 16, 6144, 30000, 1), (16, 16, 8, 6144, 700, 1), (16, 16, 16, 816, 5000, 1), (16, 8, 16, 408, 20000, 1),
         (16, 8, 8, 800, 32767, 1), (16, 16, 5, 1008, 4000, 1), (16, 16, 8, 5824, 700, 1), (8, 32, 16, 6144, 127, 1), (8, 32, 8, 2112, 60, 0),
         (8, 16, 16, 816, 127, 1), (8, 16, 8, 2048, 100, 1), (8, 16, 8, 1008, 127, 1), (16, 8, 8, 6144, 500, 1), (8, 16, 8, 6144, 100, 1)]


@pytest.mark.parametrize("bits,N,L,K,amp,use_apr", CASES)
def test_schedule_equals_oracle(port, emul, bits, N, L, K, amp, use_apr):
    rng = np.random.default_rng(K + amp)
    a_in, a_par = random_llr(rng, K + 3, amp, np.int16), random_llr(rng, K + 3, amp, np.int16)
    a_apr = random_llr(rng, K + 3, amp, np.int16) if use_apr else None
    want = port.map_win(bits, N, a_in, a_apr, a_par, K)
    got = np.zeros(K + 3, np.int16)
    assert emul.emul_map_win(bits, N, L, K, _p(a_in), _p(a_apr), _p(a_par), _p(got)) == 0
    assert (got[:K] == want).all()


def test_fast16_monitor_is_sound(port, emul):
    """Fast16 (wrapping packed arithmetic + range monitor): whenever the monitor does NOT raise the replay flag the
    output must equal the saturating oracle; across the amplitude sweep both outcomes must occur."""
    rng = np.random.default_rng(6)
    flagged = clean = 0
    for trial in range(160):
        N = (8, 16)[trial % 2]
        K = (408, 816, 1008, 2048, 6144, 5824, 512, 800)[trial % 8]
        if N == 16 and K <= 800:
            N = 8
        if K % N or K // N < 40:
            continue
        amp = int(10 ** rng.uniform(1.0, 4.3))
        aamp = min(int(amp * rng.uniform(0.5, 4)), 32767)
        a_in, a_par = random_llr(rng, K + 3, amp, np.int16), random_llr(rng, K + 3, amp, np.int16)
        a_apr = random_llr(rng, K + 3, aamp, np.int16) if trial % 3 else None
        g = int(np.abs(a_in[:K]).max()) + int(np.abs(a_par[:K]).max()) + (int(np.abs(a_apr[:K]).max()) if a_apr is not None else 0)
        want = port.map_win(16, N, a_in, a_apr, a_par, K)
        got, st = np.zeros(K + 3, np.int16), np.zeros(2, np.int32)
        f = emul.emul_map_fast16(N, K, _p(a_in), _p(a_apr), _p(a_par), _p(got), g, _p(st))
        if f:
            flagged += 1
        else:
            clean += 1
            assert (got[:K] == want).all(), (N, K, amp, aamp, g, st.tolist())
    assert flagged > 20 and clean > 20


def test_f16_kernel_arithmetic_is_sound(port, emul):
    """The arithmetic of the hot kernel k_map_f16 (factored LLR, head monitor for lane 0's known start state, tracking
    points of its passes): whenever none of its monitors raises the replay flag the output must equal the saturating
    oracle; across the amplitude sweep both outcomes must occur."""
    rng = np.random.default_rng(7)
    flagged = clean = 0
    for trial in range(200):
        N = (8, 16)[trial % 2]
        K = (408, 816, 1008, 2048, 6144, 5824, 512, 800, 6080, 1056)[trial % 10]
        if N == 16 and K <= 800:
            N = 8
        if K % N or K // N < 40:
            continue
        amp = int(10 ** rng.uniform(1.0, 4.3))
        aamp = min(int(amp * rng.uniform(0.5, 4)), 32767)
        a_in, a_par = random_llr(rng, K + 3, amp, np.int16), random_llr(rng, K + 3, amp, np.int16)
        a_apr = random_llr(rng, K + 3, aamp, np.int16) if trial % 3 else None
        g = int(np.abs(a_in[:K]).max()) + int(np.abs(a_par[:K]).max()) + (int(np.abs(a_apr[:K]).max()) if a_apr is not None else 0)
        want = port.map_win(16, N, a_in, a_apr, a_par, K)
        got = np.zeros(K + 3, np.int16)
        f = emul.emul_map_f16(N, K, _p(a_in), _p(a_apr), _p(a_par), _p(got), g)
        if f:
            flagged += 1
        else:
            clean += 1
            assert (got[:K] == want).all(), (N, K, amp, aamp, g)
    assert flagged > 20 and clean > 20


def test_glue_sub_semantics(emul):
    """srslte_vec_sub_sss wraps; srslte_vec_sub_bbb saturates in its SIMD body and wraps in its scalar tail"""
    emul.emul_glue_sub.restype = C.c_uint32
    pk = lambda lo, hi: (lo & 0xffff) | ((hi & 0xffff) << 16)
    un = lambda v: (np.int16(np.uint16(v & 0xffff)), np.int16(np.uint16(v >> 16)))
    assert un(emul.emul_glue_sub(16, pk(30000, -30000), pk(-10000, 10000), 1, 1)) == (np.int16(40000 - 65536), np.int16(-40000 + 65536))
    assert un(emul.emul_glue_sub(8, pk(100, -100), pk(-100, 100), 1, 1)) == (127, -128)
    assert un(emul.emul_glue_sub(8, pk(100, -100), pk(-100, 100), 0, 0)) == (200 - 256, -200 + 256)
    assert un(emul.emul_glue_sub(8, pk(100, -100), pk(-100, 100), 1, 0)) == (127, 56)


def test_sat8_fast_step_rules(emul):
    """Sat8F (arith.cuh): with normalised input metrics the saturations' upper halves are dead -- same integers as Sat8 for
    the backward / forward / output steps; with arbitrary input metrics the step followed by fix() is the Sat8 step"""
    for seed in range(4):
        assert emul.emul_sat8f_check(seed, 200000) == 0


# ----------------------------------------------------------------------------------------- k_ulsch_deinterleave
def test_ulsch_kernel_index_arithmetic(port):
    """The tile loops of k_ulsch_deinterleave with the index functions of srsran_b200/csrc/ulsch_core.cuh (compiled here with
    g++, tests/host_emul/emul_ulsch.cpp) against the oracle's restatement of srslte_ulsch_decode's data movement"""
    src, so = os.path.join(HERE, "host_emul", "emul_ulsch.cpp"), os.path.join(HERE, "host_emul", "libemul_ulsch.so")
    deps = [src, os.path.join(CSRC, "ulsch_core.cuh"), os.path.join(CSRC, "arith.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", so, src])
    L = C.CDLL(so)
    rng = np.random.default_rng(992)
    n = 0
    for Qm in (2, 4, 6, 8):
        for nsym in (12, 11, 10, 9):
            for rows in (1, 2, 3, 12, 63, 64, 65, 129, 300, 1200):
                if rows > 300 and Qm != 6:
                    continue
                H = rows * nsym
                for qa, qr in ((0, 0), (1, 0), (0, 1), (3, 2), (4 * rows, 4 * rows), (min(4 * rows, 37), min(4 * rows, 5)), (min(4 * rows, 6), min(4 * rows, 41))):
                    qc = int(rng.integers(0, min(H - qr, 200 if rows < 300 else 3000) + 1)) if (rows + qa) % 2 else 0
                    qbuf = rng.integers(-32768, 32768, H * Qm + 16).astype(np.int16)
                    off = (-qbuf.ctypes.data % 16) // 2 + (2 if n % 3 == 0 else 0)  # 16-byte aligned (128-bit path) or not
                    q = qbuf[off:off + H * Qm]
                    gbuf = np.full(H * Qm + 128, 777, np.int16)  # guard bands on both sides: a stray store would show
                    g = gbuf[64:64 + H * Qm]
                    uci = np.full((qa + qr + qc) * Qm + 64, 555, np.int16)
                    grid = int(rng.integers(1, 4))
                    L.emul_ulsch(_p(q), C.c_uint32(Qm), C.c_uint32(H), C.c_uint32(nsym), C.c_uint32(qa), C.c_uint32(qr), C.c_uint32(qc), _p(g), _p(uci),
                                 C.c_uint32(grid))
                    rc, g_o, ack, ri, _ = port.ulsch_deinterleave(q, Qm, nsym, qa, qr, g_fill=777)
                    assert rc == 0 and (g == g_o).all(), (Qm, nsym, rows, qa, qr, qc)
                    assert (gbuf[:64] == 777).all() and (gbuf[64 + H * Qm:] == 777).all() and (uci[(qa + qr + qc) * Qm:] == 555).all()
                    assert (uci[:qa * Qm] == ack).all() and (uci[qa * Qm:(qa + qr) * Qm] == ri).all()
                    assert (uci[(qa + qr) * Qm:(qa + qr + qc) * Qm] == g_o[:qc * Qm]).all()
                    n += 1
    assert n > 700
