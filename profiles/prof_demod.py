"""ncu driver for k_demod_descramble: 500 codewords of 15000 64QAM symbols, int16 LLRs, device-resident."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srsran_b200 as b  # noqa: E402

ctx = b.Context(0)
rng = np.random.default_rng(1)
ntb, nsym, mod, Qm = 500, 15000, 3, 6
is8 = len(sys.argv) > 1 and sys.argv[1] == "8"
if is8:
    nsym, mod, Qm = 14400, 4, 8
esz = 1 if is8 else 2
G = nsym * Qm
sym = ((rng.standard_normal((ntb, nsym)) + 1j * rng.standard_normal((ntb, nsym))) * 0.7).astype(np.complex64)
scr = b.sequence_bytes(0x12345, G)
d_sym, d_scr, d_e = ctx.device_alloc(sym.nbytes), ctx.device_alloc(len(scr) + 16), ctx.device_alloc(ntb * G * esz)
ctx.h2d(d_sym, sym)
ctx.h2d(d_scr, scr)
dm = b.make_demods(ntb)
for i in range(ntb):
    dm[i].symbols, dm[i].nof_symbols, dm[i].mod, dm[i].scramble_bytes, dm[i].e_bits = d_sym + i * nsym * 8, nsym, mod, d_scr, d_e + i * G * esz
for _ in range(4):
    ctx.demod_descramble_raw(dm, is8, b.IN_DEVICE | b.OUT_DEVICE)
ctx.timer_start()
print("drained", ctx.timer_stop_ms())
ctx.close()
