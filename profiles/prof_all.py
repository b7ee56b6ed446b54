"""One launch (at least) of EVERY kernel of the library, for a single `ncu --set full` capture:
k_prepare, k_map_fused (Fast16 8 / 16 lanes, Sat8 16 / 32 lanes, Sat16 = the exact path), k_gen_fused, k_map_gen, k_decide_crc,
k_dematch_prepare (int16 / int8), k_cb_stat, k_regroup, k_tb_finish, k_demod_descramble (int16 / int8, with and without csi), k_csi_max,
k_ulsch_deinterleave, k_map_lat + k_map_win (one subframe: the per-half-iteration path), k_enc_tb_crc, k_enc_cb."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import srsran_b200 as b  # noqa: E402

rng = np.random.default_rng(1)
ctx = b.Context(0)


def cb_batch(ncb, K, nit, dtype=np.int16):
    llr, _ = bench.make_c1(rng, ncb, K) if (K == 6144 and dtype == np.int16) else (None, None)
    if llr is None:
        llr = rng.integers(-60, 61, (ncb, 3 * K + 12)).astype(dtype)
    d_llr = ctx.device_alloc(llr.nbytes)
    d_out = ctx.device_alloc(ncb * K // 8 + 64)
    ctx.h2d(d_llr, llr)
    ctx.tdec_batch_device(d_llr, d_out, K, ncb, 3 * K + 12, 16 if dtype == np.int16 else 8, nit)


def tb_batch(wl, ntb, max_iter):
    cfg = bench.TB_CFG[wl]
    tbs, Qm, G, dt = cfg["tbs"], cfg["Qm"], cfg["G"], cfg["dtype"]
    llr, _ = bench.make_tb(rng, ntb, tbs, Qm, G, dt, cfg["amp"], cfg["sigma"])
    esz = np.dtype(dt).itemsize
    ostride = (tbs // 8 + 6 + 15) // 16 * 16
    d_llr = ctx.device_alloc(llr.nbytes)
    d_out = ctx.device_alloc(ntb * ostride)
    ctx.h2d(d_llr, llr)
    t = b.make_tbs(ntb)
    for i in range(ntb):
        t[i].e_bits, t[i].nof_e_bits, t[i].tbs, t[i].Qm, t[i].rv, t[i].softbuffer, t[i].data = d_llr + i * G * esz, G, tbs, Qm, 0, None, d_out + i * ostride
    ctx.decode_tbs(t, dt == np.int8, max_iter, flags=b.IN_DEVICE | b.OUT_DEVICE)


cb_batch(7104, 6144, 4)           # k_prepare, k_map_fused<Fast16,16> (+ the empty exact-arithmetic launch)
tb_batch("c2", 500, 8)            # k_dematch_prepare<short>, k_map_fused<Fast16,16> with decisions + CRC, k_tb_finish
tb_batch("c4", 400, 8)            # k_dematch_prepare<int8>, k_map_fused<Sat8,32>
cb_batch(4000, 2048, 3, np.int8)  # k_map_fused<Sat8,16>
ctx.set_option("fast16", 0)
cb_batch(2368, 6144, 2)           # k_map_fused<Sat16,16> doing real work (the exact path)
ctx.set_option("fast16", 1)
cb_batch(9472, 512, 3)            # k_map_fused<Fast16,8>
cb_batch(4096, 400, 6)            # k_gen_fused: the generic decoder's whole loop in one launch
ctx.set_option("gen_fused", 0)
cb_batch(4096, 40, 2)             # k_map_gen, k_decide_crc (the per-half-iteration path kept for K > 512 sessions and A/B)
ctx.set_option("gen_fused", 1)
cb_batch(13, 6144, 4)             # k_scan_fused: one subframe, time-parallel (cooperative launch)


def mixed_tb_batch(per_k):
    """all 188 sizes, single-block transport blocks of several noise levels (c3): k_cb_stat + k_regroup order the blocks of a size"""
    specs = bench.c3_specs(per_k)
    flat = np.zeros(bench.c3_elems(specs), np.int16)
    off = bench.make_c3(np.random.default_rng(1), specs, flat)
    ntb = len(specs)
    ostr = [(s_[0] // 8 + 6 + 15) // 16 * 16 for s_ in specs]
    ooff = np.concatenate([[0], np.cumsum(ostr)[:-1]])
    d_llr, d_out = ctx.device_alloc(flat.nbytes), ctx.device_alloc(int(sum(ostr)))
    ctx.h2d(d_llr, flat)
    t = b.make_tbs(ntb)
    for i in range(ntb):
        t[i].e_bits, t[i].nof_e_bits, t[i].tbs, t[i].Qm, t[i].rv, t[i].softbuffer, t[i].data = d_llr + int(off[i]) * 2, specs[i][2], specs[i][0], specs[i][1], 0, None, d_out + int(ooff[i])
    ctx.decode_tbs(t, False, 8, flags=b.IN_DEVICE | b.OUT_DEVICE)


mixed_tb_batch(16)                # k_cb_stat, k_regroup (+ every decoder class in one batch)
for dt, mod in ((np.int16, 3), (np.int8, 4)):
    n, nsym = 256, 15000
    sym = ((rng.standard_normal((n, nsym)) + 1j * rng.standard_normal((n, nsym))) * 0.7).astype(np.complex64)
    scr = b.sequence_bytes(0x12345, nsym * 8)
    d_sym, d_scr = ctx.device_alloc(sym.nbytes), ctx.device_alloc(len(scr) + 16)
    d_e = ctx.device_alloc(n * nsym * 8 * 2)
    ctx.h2d(d_sym, sym)
    ctx.h2d(d_scr, scr)
    dm = b.make_demods(n)
    Qm = b.MOD_BITS[mod]
    for i in range(n):
        dm[i].symbols, dm[i].nof_symbols, dm[i].mod, dm[i].scramble_bytes, dm[i].e_bits = d_sym + i * nsym * 8, nsym, mod, d_scr, d_e + i * nsym * Qm * np.dtype(dt).itemsize
    ctx.demod_descramble_raw(dm, dt == np.int8, b.IN_DEVICE | b.OUT_DEVICE)
# PUSCH pre-steps: k_ulsch_deinterleave<3> (100 PRB, 64QAM, ACK + RI + CQI)
ntb_ul, n_ul = 256, 1200 * 12 * 6
q_ul = rng.integers(-2000, 2000, (ntb_ul, n_ul)).astype(np.int16)
d_q, d_g = ctx.device_alloc(q_ul.nbytes), ctx.device_alloc(q_ul.nbytes)
ctx.h2d(d_q, q_ul)
ul = b.make_ulschs(ntb_ul)
for i in range(ntb_ul):
    ul[i].q_bits, ul[i].Qm, ul[i].H_prime_total, ul[i].N_pusch_symbs, ul[i].g_bits = d_q + i * n_ul * 2, 6, 14400, 12, d_g + i * n_ul * 2
    ul[i].Q_prime_ack, ul[i].Q_prime_ri, ul[i].Q_prime_cqi = 36, 20, 57
ctx.ulsch_deinterleave_raw(ul, b.IN_DEVICE | b.OUT_DEVICE)
# one subframe at a time, one launch per half-iteration: k_map_lat<Fast16,16,*> (13 blocks of K = 6144) and k_map_lat<Sat8,32,*> (one 97896-bit block)
ctx.set_option("scan", 0)
cb_batch(13, 6144, 3)
ctx.set_option("scan", 1)
tb_batch("c4", 1, 3)
# transmit mirror: k_enc_tb_crc, k_enc_cb
ntb, tbs, Qm, G = 364, 75376, 6, 90000
pay = rng.integers(0, 256, (ntb, tbs // 8), dtype=np.uint8)
ew = (G + 31) // 32 * 4
d_pay, d_eb = ctx.device_alloc(pay.nbytes), ctx.device_alloc(ntb * ew)
ctx.h2d(d_pay, pay)
en = b.make_encs(ntb)
for i in range(ntb):
    en[i].data, en[i].tbs, en[i].Qm, en[i].rv, en[i].nof_e_bits, en[i].e_bits = d_pay + i * (tbs // 8), tbs, Qm, 0, G, d_eb + i * ew
ctx.encode_tbs_raw(en, b.IN_DEVICE | b.OUT_DEVICE)
ctx.wait()
print("done")
