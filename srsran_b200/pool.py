"""Pooled vRAN receive path over the C ABI (BASELINE config 5): the subframes of many cells decoded by the GPUs
of one box.  Host-side orchestration only -- every computation is a call into libsrslte_fec_b200.so.

Sharding: a cell belongs to ONE rank (``owner_of``), so the HARQ soft buffers of its (direction, HARQ process) pairs
stay resident in the HBM of the GPU that decodes it; code blocks and subframes are independent, so there is no
collective on the data path (reference: one srslte_sch_t + srslte_softbuffer_rx_t per PHY worker and carrier,
lib/src/phy/phch/sch.c:363-570, srsenb/src/phy/cc_worker.cc).

Per rank two engines run side by side: downlink-shaped transport blocks (rate-matched e-bits as PDSCH delivers them,
decode_tb sch.c:503-570) on one, uplink subframes (descrambled q_bits -> UCI extraction + channel de-interleaving ->
decode_tb, srslte_ulsch_decode sch.c:1105-1180) on the other, so that the uplink pre-step overlaps the downlink decode.
"""
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

from . import IN_DEVICE, OUT_DEVICE, UCI_DEFERRED, B200Error, Context, Ulsch, make_tbs


def owner_of(cell: int, world: int) -> int:
    """rank that owns a cell: stable per cell, so HARQ state never moves between GPUs"""
    return cell % world


def partition(jobs, rank: int, world: int):
    """the jobs of one rank, in submission order"""
    return [j for j in jobs if owner_of(j.cell, world) == rank]


@dataclass
class Job:
    cell: int
    tti: int
    kind: str                    # "dl": llr = rate-matched e-bits; "ul": llr = descrambled q_bits of the PUSCH subframe
    pid: int                     # HARQ process
    rv: int
    new_data: bool               # first transmission of a transport block: the process's soft buffer is reset
    tbs: int
    Qm: int
    llr: object                  # int16 host array, or the address of one (pinned ring slot) with n_llr elements
    n_pusch_symbs: int = 12      # "ul" only
    q_prime: Tuple[int, int, int] = (0, 0, 0)  # "ul" only: coded symbols of ACK, RI, CQI
    n_llr: int = 0               # element count when llr is an address


@dataclass
class Result:
    ret: int = 0
    data: Optional[np.ndarray] = None
    avg_iterations: float = 0.0
    cb_noi: list = field(default_factory=list)
    cb_crc: list = field(default_factory=list)
    ack_llr: Optional[np.ndarray] = None
    ri_llr: Optional[np.ndarray] = None
    cqi_llr: Optional[np.ndarray] = None


class Batch:
    """descriptor arrays of one prepared batch (CellPool.prepare)"""

    def __init__(self):
        self.res, self.dl, self.ul, self.keep, self.resets = [], [], [], [], []
        self.t_dl = self.t_ul = self.arr_ul = None
        self.g_dev = None


class CellPool:
    """The cells one rank owns: two engines on its GPU and the device-resident HARQ soft buffers of its cells."""

    def __init__(self, rank=0, world=1, device=0, max_iterations=8, latency=None):
        """latency: None keeps the engines' own choice of kernels per batch size; 0 = throughput pooling: small batches also
        run the persistent kernel (one launch per batch instead of one per half-iteration; several pools share the GPU)"""
        self.rank, self.world, self.max_iterations = rank, world, max_iterations
        self.dl = Context(device)
        self.ul = Context(device)
        if latency is not None:
            self.dl.set_option("latency", latency)
            self.ul.set_option("latency", latency)
        self._harq = {}

    def close(self):
        for kind, sb in self._harq.values():
            (self.dl if kind == "dl" else self.ul).softbuffer_free(sb)
        self._harq = {}
        self.dl.close()
        self.ul.close()

    def owns(self, cell):
        return owner_of(cell, self.world) == self.rank

    def _softbuffer(self, job):
        key = (job.cell, job.kind, job.pid)
        if key not in self._harq:
            self._harq[key] = (job.kind, (self.dl if job.kind == "dl" else self.ul).softbuffer_create())
        return self._harq[key][1]

    def prepare(self, jobs):
        """Build the descriptor arrays of one batch of jobs (all owned by this rank; at most one job per (cell, kind, pid)).
        The batch can be run many times (bench: static grants over input rings): run() = submit() + wait()."""
        seen = set()
        for j in jobs:
            if not self.owns(j.cell):
                raise B200Error("cell %d belongs to rank %d, not %d" % (j.cell, owner_of(j.cell, self.world), self.rank))
            if (j.cell, j.kind, j.pid) in seen:
                raise B200Error("two jobs of one HARQ process in a batch: their order would matter")
            seen.add((j.cell, j.kind, j.pid))
        bt = Batch()
        bt.res = [Result() for _ in jobs]
        bt.dl = [i for i, j in enumerate(jobs) if j.kind == "dl"]
        bt.ul = [i for i, j in enumerate(jobs) if j.kind == "ul"]
        # ---- downlink-shaped blocks
        bt.t_dl = make_tbs(len(bt.dl))
        for k, i in enumerate(bt.dl):
            j = jobs[i]
            llr = j.llr if isinstance(j.llr, int) else np.ascontiguousarray(j.llr, np.int16)
            n = j.n_llr if isinstance(j.llr, int) else len(llr)
            bt.res[i].data = np.zeros(j.tbs // 8 + 8, np.uint8)
            bt.keep.append(llr)
            sb = self._softbuffer(j)
            if j.new_data:
                bt.resets.append((self.dl, sb))
            t = bt.t_dl[k]
            t.e_bits, t.nof_e_bits, t.tbs, t.Qm, t.rv, t.softbuffer, t.data = llr if isinstance(llr, int) else llr.ctypes.data, n, j.tbs, j.Qm, j.rv, sb, bt.res[i].data.ctypes.data
        # ---- uplink subframes: UCI extraction + de-interleaving into device memory, then decode from there
        bt.t_ul = make_tbs(len(bt.ul))
        if bt.ul:
            n_of = lambda j: j.n_llr if isinstance(j.llr, int) else len(j.llr)
            need = sum(n_of(jobs[i]) * 2 + 64 for i in bt.ul)
            bt.g_dev = self.ul.device_alloc(need)
            bt.arr_ul = (Ulsch * len(bt.ul))()
            off = 0
            for k, i in enumerate(bt.ul):
                j = jobs[i]
                q = j.llr if isinstance(j.llr, int) else np.ascontiguousarray(j.llr, np.int16)
                n = n_of(j)
                qa, qr, qc = j.q_prime
                H = n // j.Qm
                r = bt.res[i]
                r.ack_llr, r.ri_llr, r.cqi_llr = np.zeros(qa * j.Qm, np.int16), np.zeros(qr * j.Qm, np.int16), np.zeros(qc * j.Qm, np.int16)
                r.data = np.zeros(j.tbs // 8 + 8, np.uint8)
                bt.keep.append(q)
                sb = self._softbuffer(j)
                if j.new_data:
                    bt.resets.append((self.ul, sb))
                a = bt.arr_ul[k]
                a.q_bits, a.Qm, a.H_prime_total, a.N_pusch_symbs, a.g_bits = q if isinstance(q, int) else q.ctypes.data, j.Qm, H, j.n_pusch_symbs, bt.g_dev + off
                a.Q_prime_ack, a.Q_prime_ri, a.Q_prime_cqi = qa, qr, qc
                a.ack_llr, a.ri_llr, a.cqi_llr = r.ack_llr.ctypes.data, r.ri_llr.ctypes.data, r.cqi_llr.ctypes.data
                t = bt.t_ul[k]
                t.e_bits, t.nof_e_bits = bt.g_dev + off + qc * j.Qm * 2, (H - qr - qc) * j.Qm  # sch.c:1174-1177
                t.tbs, t.Qm, t.rv, t.softbuffer, t.data = j.tbs, j.Qm, j.rv, sb, r.data.ctypes.data
                off += (n * 2 + 63) // 64 * 64
        return bt

    def submit(self, bt):
        for ctx, sb in bt.resets:  # first transmissions: srslte_softbuffer_rx_reset (lazy: no device work)
            ctx.softbuffer_reset(sb)
        if bt.dl:
            self.dl.decode_tbs(bt.t_dl, False, self.max_iterations, submit_only=True)
        if bt.ul:
            # (the UCI LLRs come back with the data: wait() fills them, so this call only enqueues)
            self.ul.ulsch_deinterleave_raw(bt.arr_ul, OUT_DEVICE | UCI_DEFERRED)
            self.ul.decode_tbs(bt.t_ul, False, self.max_iterations, flags=IN_DEVICE, submit_only=True)

    def wait(self, bt, collect=True):
        """collect=False only waits; the results stay in the descriptor arrays bt.t_dl / bt.t_ul"""
        if bt.dl:
            self.dl.wait()
        if bt.ul:
            self.ul.wait()
        if not collect:
            return None
        for ts, idx in ((bt.t_dl, bt.dl), (bt.t_ul, bt.ul)):
            for k, i in enumerate(idx):
                t, r = ts[k], bt.res[i]
                r.ret, r.avg_iterations = t.ret, t.avg_iterations
                r.cb_noi, r.cb_crc = list(t.cb_noi[:t.nof_cb]), list(t.cb_crc[:t.nof_cb])
        return bt.res

    def release(self, bt):
        if bt.g_dev:
            self.ul.device_free(bt.g_dev)
            bt.g_dev = None

    def decode(self, jobs):
        """Decode one batch of jobs; returns a Result per job."""
        bt = self.prepare(jobs)
        self.submit(bt)
        res = self.wait(bt)
        self.release(bt)
        return res
