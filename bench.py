#!/usr/bin/env python
"""bench.py -- turbo-decoded Mbit/s on the B200 LTE turbo-decode engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c4] [--ncb N]

A "step" is one pass of the hot path over one batch of synthetic input:

  c1 (default, the configuration the metric is quoted on): NCB code blocks of K=6144, int16 LLRs in the standard
     3k+s order produced as turbodecoder_test.c:246-253 does (BPSK + AWGN, llr_s = 100*llr), exactly 4
     half-iterations each (srslte_tdec_run_all semantics, no early stop), decided bytes out.
  c2: NTB transport blocks of TBS 75376 (13 x K=5824, 64QAM-sized G=90000 e-bits, rv 0): rate de-matching +
     decode with CRC early stopping (max 8 half-iterations) + TB CRC, through the batched decode_tb entry.
  c4: same with int8 LLRs, TBS 97896 (16 x K=6144, G=115200).

`value` is whole-job decoded information Mbit/s with the inputs already resident in HBM; `e2e` is the same metric
through the C ABI with pinned HOST buffers (H2D of the LLRs and D2H of the bytes inside the timed region).
For N > 1 launch with torchrun (one rank per GPU): work is sharded by batch (weak scaling), no collective on the
data path; the timed region is bracketed by a barrier and the maximum over ranks is taken.

--impl reference times the reference's own CPU implementation (oracle/_ref, all host cores) on a bounded sample
of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W16_OPS = 82.0   # algorithmic int16 lane-ops per trellis step per half-iteration
W8_OPS = 106.0
# DRAM traffic of the dominant kernel per code block and launch (dram__bytes_read.sum + dram__bytes_write.sum of one
# `ncu --set full` capture divided by the code blocks of that launch; see profiles/README.md for the capture)
MAP_TRAFFIC_PER_CB = {"c1": {"bytes_per_cb": 92.1e3, "source": "profiles/r01_k_map_f16.metrics.txt: mean over the 4 half-iteration launches of a batch (DEC1 first 82.2, DEC2 82.3, DEC1 109.4 KB per code block without the a-posteriori plane; the last DEC2 writes it: 94.3)"}}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        """start of the timed region: wait (bounded) until nvidia-smi has produced its first sample"""
        t0 = time.perf_counter()
        while self.proc and not self.lines and time.perf_counter() - t0 < 3.0:
            time.sleep(0.05)
        self.t_mark = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        t_end = time.perf_counter()
        t_mark = getattr(self, "t_mark", 0.0)
        rows = [ln for (t, ln) in self.lines if t >= t_mark - 0.05] or [ln for (t, ln) in self.lines[-2:]]
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(n_gpus, backend=None):
    """One process per GPU.  Returns (rank, world, local_rank, barrier_fn, max_fn, sum_fn).
    backend: "nccl" on the GPU box; "gloo" (CPU tensors) is used by the world_size-2 CPU test of this plumbing."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, 0, (lambda: None), (lambda x: x), (lambda x: x)
    import torch
    import torch.distributed as dist
    backend = backend or os.environ.get("BENCH_BACKEND", "nccl")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dev = "cuda"
    else:
        dist.init_process_group(backend)
        dev = "cpu"

    def barrier():
        dist.barrier()
        if dev == "cuda":
            torch.cuda.synchronize()

    def vmax(x):
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def vsum(x):
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    return rank, world, local, barrier, vmax, vsum


def shard_seed(rank):
    """each rank decodes its own, independently generated shard of code blocks / subframes (weak scaling)"""
    return 1234 + rank


def aggregate(units_per_step_local, steps, ms_local, vmax, vsum):
    """whole-job throughput: units of all ranks over the slowest rank's device time"""
    ms = vmax(ms_local)
    total_units = vsum(units_per_step_local) * steps
    return total_units / (ms * 1e-3) / 1e6, ms, total_units


# ----------------------------------------------------------------------------------------------- workloads
def make_c1(rng, ncb, K=6144, base=64, sigma=0.9, host_out=None):
    """ncb AWGN-corrupted code words of `base` distinct random messages, int16, standard order"""
    from srsran_b200 import synth
    bits = rng.integers(0, 2, (base, K), dtype=np.uint8)
    cw = synth.turbo_encode(bits)
    stride = 3 * K + 12
    out = host_out if host_out is not None else np.zeros((ncb, stride), np.int16)
    step = 1024
    for i in range(0, ncb, step):
        n = min(step, ncb - i)
        out[i:i + n] = synth.awgn_llr(rng, cw[(np.arange(i, i + n)) % base], 100.0, sigma, np.int16)
    return out, bits


def make_tb(rng, ntb, tbs, Qm, G, dtype, amp, sigma, base=8, host_out=None):
    from srsran_b200 import synth
    data = rng.integers(0, 256, (base, tbs // 8), dtype=np.uint8)
    e = synth.encode_tbs(data, tbs, Qm, G, 0)
    out = host_out if host_out is not None else np.zeros((ntb, G), dtype)
    step = 64
    for i in range(0, ntb, step):
        n = min(step, ntb - i)
        out[i:i + n] = synth.awgn_llr(rng, e[(np.arange(i, i + n)) % base], amp, sigma, dtype)
    return out, data


TB_CFG = {"c2": dict(tbs=75376, Qm=6, G=90000, dtype=np.int16, amp=100.0, sigma=0.44, max_iter=8, C=13, K=5824),
          "c4": dict(tbs=97896, Qm=8, G=115200, dtype=np.int8, amp=20.0, sigma=0.40, max_iter=8, C=16, K=6144)}


def cpu_baseline_c1(llr, K, nof_iter, budget_s):
    """reference (or port) on the host cores over a bounded sample of the same code blocks"""
    from oracle.bindings import Port, Ref
    cores = os.cpu_count() or 1
    sample = np.ascontiguousarray(llr)
    if Ref.available():
        R = Ref()
        t1, _ = R.bench_c1(cores, sample, K, nof_iter, layout_sb=False)  # warm-up + calibration
        rep = max(1, int(budget_s / max(t1, 1e-3)))
        t_tot, _ = R.bench_c1(cores, sample, K, nof_iter, layout_sb=False, repeat=rep)
        if t_tot < 0.6 * budget_s:  # the calibration call ran cold
            rep = max(1, int(rep * budget_s / max(t_tot, 1e-3)))
            t_tot, _ = R.bench_c1(cores, sample, K, nof_iter, layout_sb=False, repeat=rep)
        n_tot = rep * len(sample)
        return {"value": n_tot * K / t_tot / 1e6, "unit": "Mbit/s", "cores": cores, "kind": "reference",
                "sample": "%d x K=%d code blocks x %d half-iterations, srslte_tdec_run_all (AUTO->avx16), %.1f s" % (n_tot, K, nof_iter, t_tot)}
    P = Port()
    h = P.tdec_new(0, True)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < budget_s:
        P.tdec_run_all(h, sample[n % len(sample)], nof_iter, K)
        n += 1
    t = time.perf_counter() - t0
    P.tdec_del(h)
    return {"value": n * K / t / 1e6, "unit": "Mbit/s", "cores": 1, "kind": "port", "sample": "%d x K=%d code blocks, scalar oracle port, %.1f s" % (n, K, t)}


def cpu_baseline_tb(llr, cfg, budget_s):
    from oracle.bindings import Port, Ref
    cores = os.cpu_count() or 1
    if Ref.available():
        R = Ref()
        t1, _, _, _ = R.bench_tb(cores, llr, cfg["tbs"], cfg["Qm"], 0, cfg["max_iter"])  # warm-up + calibration
        rep = max(1, int(budget_s / max(t1, 1e-3)))
        t_tot, _, rc, avg = R.bench_tb(cores, llr, cfg["tbs"], cfg["Qm"], 0, cfg["max_iter"], repeat=rep)
        if t_tot < 0.6 * budget_s:
            rep = max(1, int(rep * budget_s / max(t_tot, 1e-3)))
            t_tot, _, rc, avg = R.bench_tb(cores, llr, cfg["tbs"], cfg["Qm"], 0, cfg["max_iter"], repeat=rep)
        n_tot = rep * len(llr)
        it = [float(avg.mean())]
        return {"value": n_tot * cfg["tbs"] / t_tot / 1e6, "unit": "Mbit/s", "cores": cores, "kind": "reference",
                "sample": "%d TBs of %d bits, srslte_dlsch_decode2, %.2f avg half-iterations, %.1f s" % (n_tot, cfg["tbs"], float(np.mean(it)), t_tot)}
    P = Port()
    sb = P.softbuffer_new()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < budget_s:
        P.softbuffer_reset(sb)
        P.decode_tb(sb, cfg["tbs"], cfg["Qm"], 0, llr[n % len(llr)], cfg["max_iter"])
        n += 1
    t = time.perf_counter() - t0
    return {"value": n * cfg["tbs"] / t / 1e6, "unit": "Mbit/s", "cores": 1, "kind": "port", "sample": "%d TBs, scalar oracle port, %.1f s" % (n, t)}


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and rank != 0:
        return 0
    from oracle.bindings import Ref, Port
    rng = np.random.default_rng(1234)
    cores = os.cpu_count() or 1
    have_ref = Ref.available()
    if args.workload == "c1":
        K, nit = 6144, 4
        ncb = 1024 * cores if have_ref else 8
        llr, _ = make_c1(rng, ncb, K)
        units = ncb * K

        def step():
            if have_ref:
                return R.bench_c1(cores, llr, K, nit, layout_sb=False)[0]
            t0 = time.perf_counter()
            for i in range(ncb):
                P.tdec_run_all(h, llr[i], nit, K)
            return time.perf_counter() - t0
        workload = "c1: K=6144 x 4 half-iterations, int16, %d code blocks per step (bounded sample)" % ncb
    else:
        cfg = TB_CFG[args.workload]
        ntb = 64 * cores if have_ref else 1
        llr, _ = make_tb(rng, ntb, cfg["tbs"], cfg["Qm"], cfg["G"], cfg["dtype"], cfg["amp"], cfg["sigma"])
        units = ntb * cfg["tbs"]

        def step():
            if have_ref:
                return R.bench_tb(cores, llr, cfg["tbs"], cfg["Qm"], 0, cfg["max_iter"])[0]
            t0 = time.perf_counter()
            for i in range(ntb):
                P.softbuffer_reset(sb)
                P.decode_tb(sb, cfg["tbs"], cfg["Qm"], 0, llr[i], cfg["max_iter"])
            return time.perf_counter() - t0
        workload = "%s: TBS %d, %d TBs per step (bounded sample)" % (args.workload, cfg["tbs"], ntb)
    if have_ref:
        R = Ref()
    else:
        P = Port()
        h = P.tdec_new(0, True)
        sb = P.softbuffer_new()
    for _ in range(args.warmup):
        step()
    t = sum(step() for _ in range(args.steps))
    val = units * args.steps / t / 1e6
    kind = "reference" if have_ref else "port"
    used = cores if have_ref else 1
    line = {"impl": "reference", "metric": "turbo_decoded_mbps", "value": val, "unit": "Mbit/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8" if args.workload == "c4" else "int16", "data": "synthetic",
            "config": {"workload": workload, "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": val, "unit": "Mbit/s", "cores": used, "kind": kind,
                             "sample": "the reference's own AVX2 decoder (oracle/_ref) on %d host threads" % used if have_ref else "scalar oracle port, 1 thread"},
            "e2e": {"value": val, "unit": "Mbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    rank, world, local, barrier, vmax, vsum = dist_setup(args.gpus)
    import srsran_b200 as b
    from srsran_b200 import synth  # noqa: F401
    peaks, peak_src = load_peaks()
    ctx = b.Context(local)
    # engines (= streams) of this GPU: batches rotate over them, so the latency-bound tail of one batch (last wave, code
    # blocks that need all their iterations) overlaps the head of the next; the end-to-end loop also overlaps the
    # H2D copy and host-side planning of one chunk with the decode of the previous ones
    n_eng = max(2, args.engines)
    all_ctx = [ctx] + [b.Context(local) for _ in range(n_eng - 1)]
    ctx2 = all_ctx[1]
    e2e_ctx = all_ctx
    rng = np.random.default_rng(shard_seed(rank))
    sampler = ClockSampler(local)
    launches = 0
    extra = {}
    replay = [0, 0]  # (code block x half-iteration) units replayed with the exact arithmetic, units run

    if args.workload == "c1":
        K, nit = 6144, 4
        ncb = args.ncb
        stride = 3 * K + 12
        pin_in = b.PinnedArray((ncb, stride), np.int16, write_combined=args.wc_inputs)
        pin_out = b.PinnedArray((ncb, K // 8), np.uint8)
        llr, _ = make_c1(rng, ncb, K, host_out=pin_in.array)
        d_llr = ctx.device_alloc(llr.nbytes)
        d_out = ctx.device_alloc(ncb * K // 8)
        ctx.h2d(d_llr, llr)
        units_per_step = ncb * K
        algo_ops_per_step = W16_OPS * K * nit * ncb
        algo_bytes_per_step = ncb * (stride * 2 + K // 8)
        h2d_b, d2h_b = llr.nbytes, ncb * K // 8
        workload = "c1-batched: %d code blocks x K=6144 x 4 half-iterations, int16 LLRs (standard order), no early stop" % ncb

        engines = [(ctx, d_out)] + [(c, ctx.device_alloc(ncb * K // 8)) for c in all_ctx[1:]]

        def dev_submit(i):
            e, o = engines[i % n_eng]
            e.wait()  # the batch this engine ran two steps ago
            e.tdec_batch_device(d_llr, o, K, ncb, stride, 16, nit, input_sb=False, submit_only=True)

        chunks = args.e2e_chunks
        per = (ncb + chunks - 1) // chunks

        def e2e_step():
            # (the engines keep streaming across steps; e2e_drain() closes the timed region)
            cs = e2e_ctx
            for c in range(chunks):
                lo, hi = c * per, min(ncb, (c + 1) * per)
                if lo >= hi:
                    break
                e = cs[c % len(cs)]
                e.wait()
                e.tdec_batch_submit(pin_in.array[lo:hi].ctypes.data, pin_out.array[lo:hi].ctypes.data, K, hi - lo, stride, 16, nit)

        def cpu_base():
            return cpu_baseline_c1(np.array(llr[:min(ncb, 2048)]), K, nit, args.cpu_seconds)  # (a copy in ordinary memory)
    else:
        cfg = TB_CFG[args.workload]
        ntb = args.ntb
        tbs, Qm, G, dt = cfg["tbs"], cfg["Qm"], cfg["G"], cfg["dtype"]
        pin_in = b.PinnedArray((ntb, G), dt)
        ostride = (tbs // 8 + 6 + 15) // 16 * 16
        pin_out = b.PinnedArray((ntb, ostride), np.uint8)
        llr, data = make_tb(rng, ntb, tbs, Qm, G, dt, cfg["amp"], cfg["sigma"], host_out=pin_in.array)
        d_llr = ctx.device_alloc(llr.nbytes)
        d_out = ctx.device_alloc(ntb * ostride)
        ctx.h2d(d_llr, llr)
        esz = np.dtype(dt).itemsize
        tb_dev = b.make_tbs(ntb)
        for i in range(ntb):
            t = tb_dev[i]
            t.e_bits, t.nof_e_bits, t.tbs, t.Qm, t.rv, t.softbuffer, t.data = d_llr + i * G * esz, G, tbs, Qm, 0, None, d_out + i * ostride
        units_per_step = ntb * tbs
        h2d_b, d2h_b = llr.nbytes, ntb * (tbs // 8 + 6)
        workload = "%s: %d TBs x TBS %d (%d x K=%d), %s e-bits, rate de-matching + <=%d half-iterations with CRC early stop" % (
            args.workload, ntb, tbs, cfg["C"], cfg["K"], np.dtype(dt).name, cfg["max_iter"])

        engines = [(ctx, tb_dev)]
        for c in all_ctx[1:]:
            d_o = ctx.device_alloc(ntb * ostride)
            tbd = b.make_tbs(ntb)
            for i in range(ntb):
                t = tbd[i]
                t.e_bits, t.nof_e_bits, t.tbs, t.Qm, t.rv, t.softbuffer, t.data = d_llr + i * G * esz, G, tbs, Qm, 0, None, d_o + i * ostride
            engines.append((c, tbd))

        def dev_submit(i):
            e, t = engines[i % n_eng]
            e.wait()
            e.decode_tbs(t, dt == np.int8, cfg["max_iter"], flags=b.IN_DEVICE | b.OUT_DEVICE, submit_only=True)

        # end to end: the TTIs of a step go down in 4 chunks that alternate between the two engines, so the H2D copy
        # of one chunk overlaps the decode of the previous one
        n_chunks = args.e2e_chunks
        per = (ntb + n_chunks - 1) // n_chunks
        host_chunks = []
        for c in range(n_chunks):
            lo, hi = c * per, min(ntb, (c + 1) * per)
            if lo >= hi:
                break
            arr = b.make_tbs(hi - lo)
            for i in range(lo, hi):
                t = arr[i - lo]
                t.e_bits, t.nof_e_bits, t.tbs, t.Qm, t.rv, t.softbuffer, t.data = pin_in.ptr + i * G * esz, G, tbs, Qm, 0, None, pin_out.ptr + i * ostride
            host_chunks.append(arr)

        def e2e_step():
            cs = e2e_ctx
            for c, arr in enumerate(host_chunks):
                e = cs[c % len(cs)]
                e.wait()
                e.decode_tbs(arr, dt == np.int8, cfg["max_iter"], flags=0, submit_only=True)

        def cpu_base():
            return cpu_baseline_tb(llr[:min(ntb, 64)], cfg, args.cpu_seconds)

    # ---- device-resident throughput (`value`): two engines (streams) of this GPU alternate batches, so the tail of
    #      one batch (last wave / code blocks that need all their iterations) overlaps the head of the next
    def stats(e):
        replay[0] += e.last_replayed()
        replay[1] += e.last_half_iterations()
        return e.last_launches(), e.last_map_ms(), e.last_map_launches()

    sampler.start()
    for i in range(max(args.warmup, n_eng)):
        dev_submit(i)
    for c in all_ctx:
        c.wait()
    replay[0] = replay[1] = 0
    barrier()
    sampler.mark()
    ctx.timer_start()
    map_ms, map_launches = 0.0, 0
    for i in range(args.steps):
        e = engines[i % n_eng][0]
        if i >= n_eng:
            e.wait()
            l, mm, ml = stats(e)
            launches += l
            map_ms += mm
            map_launches += ml
        dev_submit(i)
    for i in range(max(0, args.steps - n_eng), args.steps): # the batches still in flight, oldest first
        e = engines[i % n_eng][0]
        e.wait()
        l, mm, ml = stats(e)
        launches += l
        map_ms += mm
        map_launches += ml
    ms = ctx.timer_stop_ms()
    extra["exact_replay_fraction"] = replay[0] / max(1, replay[1])
    barrier()
    value, ms, total_units = aggregate(units_per_step, args.steps, ms, vmax, vsum)

    # ---- end to end through the C ABI with host buffers
    def e2e_drain():
        for e in e2e_ctx:
            e.wait()

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    e2e_drain()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_drain()
    e2e_s = vmax(time.perf_counter() - t0)
    clocks = sampler.stop()  # samples cover the device-resident and the end-to-end timed regions
    barrier()
    e2e_value = total_units / e2e_s / 1e6

    # ---- per-TTI latency (second half of the metric): ONE subframe's worth of work submitted from host buffers,
    #      submit -> results on the host, back to back on an otherwise idle GPU
    lat = []
    n_lat = 0 if args.no_latency else 300
    if n_lat == 0:
        tti_desc = "skipped (--no-latency)"
    elif args.workload == "c1":
        tti_cb = 13
        tti_desc = "%d code blocks x K=6144 x 4 half-iterations (one 75 kbit subframe), host buffers in and out" % tti_cb
        for i in range(n_lat + 20):
            lo = (i * tti_cb) % (ncb - tti_cb)
            t0 = time.perf_counter()
            ctx.tdec_batch_submit(pin_in.array[lo:lo + tti_cb].ctypes.data, pin_out.array[lo:lo + tti_cb].ctypes.data, K, tti_cb, stride, 16, nit)
            ctx.wait()
            lat.append(time.perf_counter() - t0)
    else:
        tti_desc = "1 transport block of %d bits (%d code blocks), <=%d half-iterations with CRC early stop, host buffers in and out" % (tbs, cfg["C"], cfg["max_iter"])
        one = b.make_tbs(1)
        for i in range(n_lat + 20):
            k = i % ntb
            t = one[0]
            t.e_bits, t.nof_e_bits, t.tbs, t.Qm, t.rv, t.softbuffer, t.data = pin_in.ptr + k * G * esz, G, tbs, Qm, 0, None, pin_out.ptr + k * ostride
            t0 = time.perf_counter()
            ctx.decode_tbs(one, dt == np.int8, cfg["max_iter"], flags=0)
            lat.append(time.perf_counter() - t0)
    if n_lat:
        lat = np.sort(np.array(lat[20:])) * 1e6
        latency = {"tti": tti_desc, "p50_us": float(lat[len(lat) // 2]), "p99_us": float(lat[int(len(lat) * 0.99)]), "max_us": float(lat[-1]), "n": int(len(lat))}
    else:
        latency = {"tti": tti_desc, "p50_us": None, "p99_us": None, "max_us": None, "n": 0}
    barrier()

    if args.workload != "c1":
        avg_it = float(np.mean([tb_dev[i].avg_iterations for i in range(ntb)]))
        K, nit = cfg["K"], avg_it
        algo_ops_per_step = (W8_OPS if cfg["dtype"] == np.int8 else W16_OPS) * K * avg_it * cfg["C"] * ntb
        algo_bytes_per_step = ntb * (G * esz + tbs // 8)
        n_ok = sum(1 for i in range(ntb) if tb_dev[i].ret == 0)
        extra["avg_half_iterations"] = avg_it
        extra["tb_ok_fraction"] = n_ok / ntb

        # ---- front end of the path: soft demodulation + descrambling of the same subframes' symbols on
        #      the device (k_demod_descramble, HBM-bound): device-resident symbols in, LLRs out, CUDA-event timed
        mod = {6: 3, 8: 4}[Qm]
        nsym = G // Qm
        sym = ((rng.standard_normal((ntb, nsym)) + 1j * rng.standard_normal((ntb, nsym))) * 0.7).astype(np.complex64)
        scr = b.sequence_bytes(0x12345, G)
        d_sym = ctx.device_alloc(sym.nbytes)
        d_scr = ctx.device_alloc(len(scr) + 16)
        d_e = ctx.device_alloc(ntb * G * esz)
        ctx.h2d(d_sym, sym)
        ctx.h2d(d_scr, scr)
        dm = b.make_demods(ntb)
        for i in range(ntb):
            dm[i].symbols, dm[i].nof_symbols, dm[i].mod, dm[i].scramble_bytes, dm[i].e_bits = d_sym + i * nsym * 8, nsym, mod, d_scr, d_e + i * G * esz
        for _ in range(3):
            ctx.demod_descramble_raw(dm, dt == np.int8, b.IN_DEVICE | b.OUT_DEVICE)
        reps = 10
        ctx.timer_start()
        for _ in range(reps):
            ctx.demod_descramble_raw(dm, dt == np.int8, b.IN_DEVICE | b.OUT_DEVICE)
        fe_ms = ctx.timer_stop_ms() / reps
        fe_bytes = ntb * (nsym * 8 + G * esz + G / 8)
        extra["front_end"] = {"kernel": "k_demod_descramble", "what": "%d codewords x %d symbols, %s, %s LLRs, device-resident" % (
                                  ntb, nsym, {3: "64QAM", 4: "256QAM"}[mod], np.dtype(dt).name),
                              "ms": fe_ms, "gsym_per_s": ntb * nsym / (fe_ms * 1e-3) / 1e9,
                              "hbm": {"achieved_gbs": fe_bytes / (fe_ms * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                                      "frac": fe_bytes / (fe_ms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0)}}
        for p_ in (d_sym, d_scr, d_e):
            ctx.device_free(p_)

        # ---- PUSCH pre-steps: a 100-PRB 64QAM uplink subframe per transport block -- ACK / RI / CQI LLR
        #      extraction + channel de-interleaving (k_ulsch_deinterleave, HBM-bound), device-resident, CUDA-event timed
        if dt == np.int16:
            ul_rows, ul_cols, ul_Qm, ul_q = 1200, 12, 6, (36, 20, 57)
            ul_n = ul_rows * ul_cols * ul_Qm
            qll = rng.integers(-2000, 2000, (min(ntb, 256), ul_n)).astype(np.int16)
            d_ulq = ctx.device_alloc(ntb * ul_n * 2)
            d_ulg = ctx.device_alloc(ntb * ul_n * 2)
            for i in range(0, ntb, len(qll)):
                k = min(len(qll), ntb - i)
                ctx.h2d(d_ulq + i * ul_n * 2, qll[:k])
            ul = b.make_ulschs(ntb)
            for i in range(ntb):
                ul[i].q_bits, ul[i].Qm, ul[i].H_prime_total, ul[i].N_pusch_symbs, ul[i].g_bits = d_ulq + i * ul_n * 2, ul_Qm, ul_rows * ul_cols, ul_cols, d_ulg + i * ul_n * 2
                ul[i].Q_prime_ack, ul[i].Q_prime_ri, ul[i].Q_prime_cqi = ul_q
            for _ in range(3):
                ctx.ulsch_deinterleave_raw(ul, b.IN_DEVICE | b.OUT_DEVICE)
            ctx.timer_start()
            for _ in range(10):
                ctx.ulsch_deinterleave_raw(ul, b.IN_DEVICE | b.OUT_DEVICE)
            ul_ms = ctx.timer_stop_ms() / 10
            ul_bytes = ntb * (2 * ul_n * 2 - ul_q[1] * ul_Qm * 2)
            extra["ul_pre"] = {"kernel": "k_ulsch_deinterleave", "what": "%d subframes x %d x %d symbols, 64QAM, Q'(ack, ri, cqi) = %s, device-resident" % (
                                   ntb, ul_rows, ul_cols, list(ul_q)),
                               "ms": ul_ms, "gllr_per_s": ntb * ul_n / (ul_ms * 1e-3) / 1e9,
                               "hbm": {"achieved_gbs": ul_bytes / (ul_ms * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                                       "frac": ul_bytes / (ul_ms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0)}}
            ctx.device_free(d_ulq)
            ctx.device_free(d_ulg)

        # ---- transmit mirror of the same transport blocks: payload bytes -> packed e-bits, device-resident
        pay = rng.integers(0, 256, (ntb, tbs // 8), dtype=np.uint8)
        ew = (G + 31) // 32 * 4
        d_pay = ctx.device_alloc(pay.nbytes)
        d_eb = ctx.device_alloc(ntb * ew)
        ctx.h2d(d_pay, pay)
        en = b.make_encs(ntb)
        for i in range(ntb):
            en[i].data, en[i].tbs, en[i].Qm, en[i].rv, en[i].nof_e_bits, en[i].e_bits = d_pay + i * (tbs // 8), tbs, Qm, 0, G, d_eb + i * ew
        for _ in range(3):
            ctx.encode_tbs_raw(en, b.IN_DEVICE | b.OUT_DEVICE)
        ctx.timer_start()
        for _ in range(reps):
            ctx.encode_tbs_raw(en, b.IN_DEVICE | b.OUT_DEVICE)
        tx_ms = ctx.timer_stop_ms() / reps
        extra["tx_mirror"] = {"kernel": "k_enc_tb_crc + k_enc_cb", "what": "%d transport blocks of %d bits -> %d e-bits each, device-resident" % (ntb, tbs, G),
                              "ms": tx_ms, "encoded_mbps": ntb * tbs / (tx_ms * 1e-3) / 1e6}
        ctx.device_free(d_pay)
        ctx.device_free(d_eb)

    # ---- roofline of the dominant kernel (k_map_win): integer-ALU issue bound.  The kernel is timed on its own here
    #      (one engine, batches back to back, CUDA events around every launch on the launching stream).
    rsteps = max(1, min(5, args.steps))
    ctx.timer_start()
    r_map_ms, r_map_launches, r_gpu_ms = 0.0, 0, 0.0
    for i in range(rsteps):
        dev_submit(0)
        ctx.wait()
        r_map_ms += ctx.last_map_ms()
        r_gpu_ms += ctx.last_gpu_ms()
        r_map_launches += ctx.last_map_launches()
    ms_single = ctx.timer_stop_ms() / rsteps
    # peak: issue rate of the packed instructions the kernel is made of, each measured on its own (VIADD.16x2, VIMNMX.S16x2,
    # VIADDMNMX.S16x2, VIMNMX3.S16x2 all sustain 0.5 warp-instructions/clk/SM sub-partition on B200: 16-lane integer pipe)
    probes = [ctx.alu_probe(op) for op in range(4)]
    probe = min(probes)                # packed int16x2 instructions / s
    probe_sat = ctx.alu_probe(4)
    peak_lane_ops = 2.0 * probe        # one algorithmic operation per int16 lane and instruction
    # the dominant kernel is the persistent k_map_fused: ONE launch runs all half-iterations of a batch (a second, exact-
    # arithmetic launch follows and returns at once unless the range monitor parked blocks); `achieved` is the batch's
    # algorithmic work over the event-timed duration of those launches
    map_s_per_launch = (r_map_ms * 1e-3) / max(1, rsteps)
    ach_lane_ops = (algo_ops_per_step / map_s_per_launch) if r_map_ms > 0 else 0.0
    map_ms, map_launches = r_map_ms, r_map_launches
    traffic = MAP_TRAFFIC_PER_CB.get(args.workload)
    roofline = {"bound": "int_alu", "achieved": ach_lane_ops / 1e12, "peak": peak_lane_ops / 1e12, "unit": "Tlane-op/s (int16)",
                "frac": ach_lane_ops / peak_lane_ops if peak_lane_ops else None,
                "traffic": None if traffic is None else traffic["bytes_per_cb"] * (units_per_step / K if args.workload == "c1" else cfg["C"] * ntb),
                "traffic_source": None if traffic is None else traffic["source"],
                "kernel": "k_map_fused", "launches": map_launches, "avg_launch_ms": 1e3 * map_s_per_launch,
                # share of the batch's device time (first kernel -> last kernel, CUDA events) spent in the MAP launches
                "map_share_of_step": r_map_ms / r_gpu_ms if r_gpu_ms else None, "batch_device_ms": r_gpu_ms / rsteps,
                "single_stream_mbps": units_per_step / (ms_single * 1e-3) / 1e6,
                "peak_source": "live micro-benchmark (k_alu_probe): min over VIADD.16x2 / VIMNMX.S16x2 / VIADDMNMX.S16x2 / VIMNMX3.S16x2 of the packed "
                               "instruction rate, x2 int16 lanes; algorithmic work = 82 (int8: 106) lane-ops per trellis step and half-iteration",
                "probe_packed_tops": [x / 1e12 for x in probes],
                "peak_with_saturating_emulation": 2.0 * probe_sat / 1e12,
                "hbm": {"achieved_gbs": algo_bytes_per_step / ((ms / args.steps) * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                        "frac": algo_bytes_per_step / ((ms / args.steps) * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0), "peak_source": peak_src}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_base()
        except Exception as ex:  # the baseline is reporting only; never fail the GPU measurement on it
            cpu = {"value": None, "unit": "Mbit/s", "cores": 0, "kind": "unavailable", "sample": str(ex)}

    if rank == 0:
        line = {"metric": "turbo_decoded_mbps", "value": value, "unit": "Mbit/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int8" if args.workload == "c4" else "int16", "data": "synthetic",
                "config": {"workload": workload, "l2": "inputs (%.0f MB per GPU per step) larger than the 126 MB L2" % (h2d_b / 1e6),
                           "parallelism": "batch-sharded x%d, no collective" % world, "engines_per_gpu": n_eng, "e2e_chunks_per_step": args.e2e_chunks},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "Mbit/s", "h2d_bytes_per_step": int(h2d_b * world), "d2h_bytes_per_step": int(d2h_b * world)},
                "gpu_launches": int(launches * world), "latency": latency, "roofline": roofline}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------------------------- pooled cells (config 5)
def run_c5(args):
    """BASELINE config 5: the subframes of `--cells` cells, one downlink-shaped transport block (C2 shape) and one uplink
    subframe (100 PRB 64QAM with ACK, RI and CQI) per cell and TTI, sharded over the ranks by cell (HARQ affinity),
    decoded through srsran_b200.pool.CellPool from pinned host buffers.  A tenth of the downlink HARQ processes
    alternate between a first transmission that fails and the retransmission that combines with it."""
    rank, world, local, barrier, vmax, vsum = dist_setup(args.gpus)
    import srsran_b200 as b
    from srsran_b200 import synth
    from srsran_b200.pool import CellPool, Job, owner_of
    rng = np.random.default_rng(shard_seed(rank))
    # two pools per GPU, each with half of the rank's cells: while one pool's batch decodes, the other's is copied and planned
    pools = [CellPool(rank, world, local) for _ in range(2)]
    cells = [c for c in range(args.cells) if owner_of(c, world) == rank]
    tbs, Qm, G = 75376, 6, 90000
    rows, nsym, qp = 1200, 12, (36, 20, 57)
    n_q = rows * nsym * Qm
    G_ul = (rows * nsym - qp[1] - qp[2]) * Qm
    base = 8
    data = rng.integers(0, 256, (base, tbs // 8), dtype=np.uint8)
    pin_dl = b.PinnedArray((base, G), np.int16)
    pin_hard = b.PinnedArray((2, G), np.int16)
    pin_ul = b.PinnedArray((base, n_q), np.int16)
    pin_dl.array[:] = synth.awgn_llr(rng, synth.encode_tbs(data, tbs, Qm, G, 0), 100.0, 0.40, np.int16)
    pin_hard.array[0] = synth.awgn_llr(rng, synth.encode_tbs(data[:1], tbs, Qm, G, 0), 100.0, 0.56, np.int16)[0]
    pin_hard.array[1] = synth.awgn_llr(rng, synth.encode_tbs(data[:1], tbs, Qm, G, 2), 100.0, 0.56, np.int16)[0]
    g_tx = np.concatenate([rng.integers(0, 2, (base, qp[2] * Qm), dtype=np.uint8), synth.encode_tbs(data, tbs, Qm, G_ul, 0)], axis=1)
    pin_ul.array[:] = synth.awgn_llr(rng, synth.ul_interleave(rng, g_tx, Qm, rows, nsym, qp[0], qp[1]), 100.0, 0.30, np.int16)
    n_pid = 8
    batches = []  # [pool][phase] -> (jobs, prepared batch)
    for pi, pool in enumerate(pools):
        per_phase = []
        for phase in range(2):  # the hard processes alternate: first transmission (fails) / retransmission (combines)
            jobs = []
            for c in cells[pi::2]:
                for pid in range(n_pid):
                    k = (c * n_pid + pid) % base
                    if (c * n_pid + pid) % 10 == 3:
                        jobs.append(Job(c, pid, "dl", pid, 2 * phase, phase == 0, tbs, Qm, pin_hard.ptr + phase * G * 2, n_llr=G))
                    else:
                        jobs.append(Job(c, pid, "dl", pid, 0, True, tbs, Qm, pin_dl.ptr + k * G * 2, n_llr=G))
                    jobs.append(Job(c, pid, "ul", pid, 0, True, tbs, Qm, pin_ul.ptr + k * n_q * 2, nsym, qp, n_llr=n_q))
            per_phase.append((jobs, pool.prepare(jobs)))
        batches.append(per_phase)
    sampler = ClockSampler(local)
    sampler.start()
    in_flight = [None, None]
    tally = [0, 0]  # decoded bits, kernel launches

    def retire(pi):
        bt = in_flight[pi]
        if bt is None:
            return
        pools[pi].wait(bt, collect=False)
        tally[0] += tbs * (sum(1 for k in range(len(bt.dl)) if bt.t_dl[k].ret == 0) + sum(1 for k in range(len(bt.ul)) if bt.t_ul[k].ret == 0))
        tally[1] += pools[pi].dl.last_launches() + pools[pi].ul.last_launches()
        in_flight[pi] = None

    def step(i):
        for pi in range(2):
            retire(pi)
            if batches[pi][0][0]:
                in_flight[pi] = batches[pi][i % 2][1]
                pools[pi].submit(in_flight[pi])

    for i in range(2 * max(1, args.warmup // 2)):
        step(i)
    retire(0)
    retire(1)
    barrier()
    sampler.mark()
    tally[0] = tally[1] = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    retire(0)
    retire(1)
    bits, launches = tally
    dt = vmax(time.perf_counter() - t0)
    clocks = sampler.stop()
    barrier()
    total = vsum(bits)
    my_jobs = batches[0][0][0] + batches[1][0][0]
    n_jobs = vsum(len(my_jobs))
    if rank == 0:
        val = total / dt / 1e6
        h2d = sum((j.n_llr * 2) for j in my_jobs)
        line = {"metric": "turbo_decoded_mbps", "value": val, "unit": "Mbit/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int16",
                "data": "synthetic",
                "config": {"workload": "c5-pooled: %d cells x %d TTIs per step (one 75376-bit downlink-shaped block + one 100-PRB 64QAM uplink subframe with "
                                       "ACK/RI/CQI per cell and TTI), cells sharded over %d rank(s), device-resident HARQ soft buffers, 10%% of the downlink "
                                       "processes alternate failing first transmission / combining retransmission" % (args.cells, n_pid, world),
                           "l2": "inputs of a step (%d MB over all ranks) larger than the 126 MB L2" % (n_jobs * G * 2 // 1000000),
                           "parallelism": "cells sharded x%d, no collective" % world, "note": "host-fed only: value == e2e (wall clock, max over ranks)"},
                "clocks": clocks, "e2e": {"value": val, "unit": "Mbit/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": len(my_jobs) * (tbs // 8 + 6)},
                "gpu_launches": int(launches), "decoded_tb_fraction": total / (n_jobs * tbs * args.steps)}
        print(json.dumps(line))
    for pi, pool in enumerate(pools):
        for _, bt in batches[pi]:
            pool.release(bt)
        pool.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1", choices=["c1", "c2", "c4", "c5"])
    ap.add_argument("--cells", type=int, default=20, help="cells of the pooled workload (c5), sharded over the ranks")
    ap.add_argument("--ncb", type=int, default=18944, help="code blocks per step per GPU (c1); 18944 = 4 full waves of 148 CTAs x 32 blocks")
    ap.add_argument("--ntb", type=int, default=1000, help="transport blocks per step per GPU (c2/c4)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline budget")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the per-TTI latency loop (keeps profiler launch lists short)")
    ap.add_argument("--engines", type=int, default=0,
                    help="engines (streams) per GPU the batches / end-to-end chunks rotate over (default: 4 for c1, 6 for the early-stop workloads "
                         "whose last half-iterations run nearly empty and overlap with other batches)")
    ap.add_argument("--wc-inputs", type=int, default=0, help="c1: allocate the host LLR staging buffer write-combined (srslte_b200_host_alloc_wc)")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="chunks one end-to-end step is split into")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.engines <= 0:
        args.engines = 4 if args.workload == "c1" else 6
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "c5":
        return run_c5(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
