"""ncu driver for k_map_lat: one subframe (13 code blocks, K=6144, 4 half-iterations) from host buffers, a few times."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srsran_b200 as b  # noqa: E402
from srsran_b200 import synth  # noqa: E402

ctx = b.Context(0)
rng = np.random.default_rng(1)
K, nit, ncb = 6144, 4, 13
stride = 3 * K + 12
bits = rng.integers(0, 2, (ncb, K), dtype=np.uint8)
llr = synth.awgn_llr(rng, synth.turbo_encode(bits), 100.0, 0.9, np.int16)
for _ in range(4):
    out = ctx.tdec_batch(np.ascontiguousarray(llr), K, nit, input_sb=False)
print("bit errors", int((np.unpackbits(out, axis=1) != bits).sum()))
ctx.close()
