"""ncu driver for k_ulsch_deinterleave: 500 uplink subframes (100 PRB, 64QAM, ACK + RI + CQI), device-resident."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srsran_b200 as b  # noqa: E402

ctx = b.Context(0)
rng = np.random.default_rng(1)
ntb, n = 500, 1200 * 12 * 6
q = rng.integers(-2000, 2000, (ntb, n)).astype(np.int16)
dq, dg = ctx.device_alloc(q.nbytes), ctx.device_alloc(q.nbytes)
ctx.h2d(dq, q)
ul = b.make_ulschs(ntb)
for i in range(ntb):
    ul[i].q_bits, ul[i].Qm, ul[i].H_prime_total, ul[i].N_pusch_symbs, ul[i].g_bits = dq + i * n * 2, 6, 14400, 12, dg + i * n * 2
    ul[i].Q_prime_ack, ul[i].Q_prime_ri, ul[i].Q_prime_cqi = 36, 20, 57
for _ in range(4):
    ctx.ulsch_deinterleave_raw(ul, b.IN_DEVICE | b.OUT_DEVICE)
ctx.timer_start()
print("drained", ctx.timer_stop_ms())
ctx.close()
