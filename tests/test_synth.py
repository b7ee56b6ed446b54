"""the numpy transmit chain used to synthesise benchmark inputs (srsran_b200/synth.py) against the oracle's encoder chain"""
import numpy as np

from srsran_b200 import build, synth


def test_synth_matches_oracle(port):
    build.build()
    rng = np.random.default_rng(3)
    for K in (40, 504, 6144):
        bits = rng.integers(0, 2, (3, K), dtype=np.uint8)
        cw = synth.turbo_encode(bits)
        for i in range(3):
            assert (cw[i] == port.tcod_encode(bits[i])).all()
    for tbs, Qm, G in ((75376, 6, 90000), (6120, 2, 14400), (15264, 4, 20000)):
        d = rng.integers(0, 256, (2, tbs // 8), dtype=np.uint8)
        for rv in (0, 2):
            e = synth.encode_tbs(d, tbs, Qm, G, rv)
            for i in range(2):
                assert (e[i] == port.encode_tb(tbs, Qm, rv, G, d[i])).all()
    for poly in (synth.CRC24A, synth.CRC24B):
        x = rng.integers(0, 256, (4, 100), dtype=np.uint8)
        assert synth.crc24(x, poly).tolist() == [port.crc_bytes(poly, 24, x[i]) for i in range(4)]
