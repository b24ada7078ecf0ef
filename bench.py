#!/usr/bin/env python
"""bench.py — the measured hot path (BASELINE.json metric) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-rows]

Workload (config.workload "cfg5"): BASELINE.json configs[4] — f32 X[262144,16384] (2^32 elements, 17.18 GB),
sharded over the OUTER axis across the N GPUs (rank r owns rows [r·262144/N, (r+1)·262144/N)): STRONG scaling,
the total size is fixed.  It is the largest configuration that fits one GPU and the only one whose reductions
cross the shard axis, i.e. the one that exercises the exchange.  One step = X.sum(), X.mean(),
X.sum(axis 0) through the public API (`hptb_reduce_sharded`): every rank reduces its shard to one accumulator
per output and the k accumulators are exchanged over NVLink peer memory INSIDE the reduce kernel's epilogue
(xchg.cuh; NCCL all-gather when peer memory is unavailable), so a step is three launches per rank.
Algorithmic bytes per step: 3 · 17,179,869,184 read + 4 + 4 + 65,536 written.
`value` = those bytes ÷ max-over-ranks device time (GB/s), shards resident in HBM.
`e2e`   = the same step with HOST buffers: every step uploads the rank's shard from pinned host memory
          (17.18/N GB over PCIe), runs the three reductions and reads the three results back.
`rows`  = the other four BASELINE configurations (device-resident, per rank — they need no exchange).
`--impl reference` times the CPU restatement of Hpt's path (oracle/oracle_cpu.cpp, OpenMP, all host threads) on
a bounded row block of the same tensor; Hpt itself is Rust and cannot be built in this image.
"""
import argparse
import ctypes
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ROWS, COLS = 262144, 16384
IN_BYTES = ROWS * COLS * 4                        # 17,179,869,184
BYTES_STEP = 3 * IN_BYTES + 4 + 4 + COLS * 4      # three passes + the three results
ELEMS_STEP = 3 * ROWS * COLS
METRIC = "hbm_gbs_sharded_sum_mean_sum0"
# one dict, printed verbatim by BOTH arms (the driver compares them)
CONFIG = {
    "workload": "cfg5: f32 [262144,16384] sharded over the outer axis: sum(), mean(), sum(axis 0)",
    "shard": "strong scaling: rank r of N owns rows [r*262144/N, (r+1)*262144/N); per-output accumulators are "
             "exchanged over NVLink peer memory inside the reduce kernel's epilogue (NCCL all-gather fallback)",
    "l2": "every pass streams 17.18/N GB (>= 2.1 GB), far above the 126 MB L2: no flush needed",
}
CPU_SAMPLE_ROWS_REF = 32768     # --impl reference: [32768,16384] f32 = 2.1 GB per pass (well beyond any last-level cache)
CPU_SAMPLE_ROWS_BASE = 32768    # cpu_baseline of our arm: [32768,16384] = 2.1 GB per pass (an 8-GPU shard)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs: torch copy, read+write bytes)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------------
# CPU restatement (oracle/oracle_cpu.cpp): cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def load_cpu_port():
    path = os.path.join(ROOT, "oracle", "_build", "liboracle_cpu.so")
    if not os.path.exists(path):
        import build as _b
        _b.build_oracle()
    L = ctypes.CDLL(path)
    vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    L.orc_sum_f32_all.argtypes = [vp, i64]
    L.orc_sum_f32_all.restype = ctypes.c_float
    L.orc_sum_f32_cols.argtypes = [vp, i64, i64, vp]
    L.orc_num_threads.restype = ci
    L.orc_set_num_threads.argtypes = [ci]
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm uses every host thread it may run on
    L.orc_set_num_threads(host_threads())
    return L


def time_cpu(steps, warmup, sample_rows):
    """The three reductions of the workload on a [sample_rows, 16384] row block, on the host: Hpt's CPU path sums in
    the output dtype (f32), mean = that sum ÷ n (common_reduce.rs:352-380), sum(axis 0) walks rows outermost."""
    import numpy as np
    L = load_cpu_port()
    rng = np.random.default_rng(1234 + 5)
    x = rng.standard_normal((sample_rows, COLS), dtype=np.float32)
    oc = np.empty(COLS, np.float32)
    n = sample_rows * COLS

    def step():
        s = L.orc_sum_f32_all(x.ctypes.data, n)
        m = L.orc_sum_f32_all(x.ctypes.data, n) / n
        L.orc_sum_f32_cols(x.ctypes.data, sample_rows, COLS, oc.ctypes.data)
        return s, m

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        s, m = step()
    dt = (time.perf_counter() - t0) / steps
    ref = float(x.sum(dtype=np.float64))
    assert abs(s - ref) <= 1e-6 * math.log2(n) * float(np.abs(x).sum(dtype=np.float64)), "CPU port: sum mismatch"
    nbytes = 3 * n * 4 + 4 + 4 + COLS * 4
    return nbytes / dt / 1e9, dt, L.orc_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    gbs, dt, cores = time_cpu(steps, warmup, CPU_SAMPLE_ROWS_REF)
    line = {"impl": "reference", "metric": METRIC, "value": round(gbs, 3), "unit": "GB/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
            "cpu_baseline": {"value": round(gbs, 3), "unit": "GB/s", "cores": cores, "kind": "port",
                             "sample": f"rows [0,{CPU_SAMPLE_ROWS_REF}) of the tensor ([{CPU_SAMPLE_ROWS_REF},16384] f32, 2.1 GB per pass), "
                                       f"{steps} timed steps of sum(), mean(), sum(axis 0) by the C++/OpenMP restatement of Hpt's CPU path "
                                       "(oracle/oracle_cpu.cpp); Hpt's Rust CPU path cannot be built here"},
            "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def mem_available_bytes():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) * 1024
    except Exception:
        pass
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from ctypes import byref, c_int32, c_void_p

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: hpt_b200 has no CPU fallback")
    if ROWS % world:
        raise SystemExit(f"--gpus {world} does not divide {ROWS} rows")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import hpt_b200 as hb
    from hpt_b200 import _ffi

    stream = torch.cuda.current_stream()
    hb.set_stream(stream.cuda_stream)
    ctx = hb.context(local)
    T = hb.Tensor
    rows_local = ROWS // world
    row0 = rank * rows_local

    # this rank's shard, generated on the device (seeded per shard; a 17 GB host tensor plus an f64 reference would not
    # fit beside it in host memory)
    g = torch.Generator(device=dev).manual_seed(1234 + 5 + 1000 * rank)
    big = torch.empty((rows_local, COLS), device=dev, dtype=torch.float32)
    for r0 in range(0, rows_local, 16384):
        big[r0:r0 + 16384].normal_(generator=g)
    X = T.from_device_ptr(big.data_ptr(), hb.F32, (rows_local, COLS), device=local, keepalive=big)

    # one NCCL rank per process; hptb_reduce_sharded at every N (a 1-rank comm makes it the plain local reduction)
    comm, exchange = None, "none (single GPU)"
    try:
        comm = hb.Comm.from_torch_distributed(ctx) if world > 1 else hb.Comm(ctx, 1, 0, hb.Comm.unique_id())
    except hb.HptError:
        if world > 1:
            raise
    if world > 1:
        exchange = ("peer-memory mailboxes, fused into the reduce kernel epilogue (one launch per rank and op)"
                    if hb.lib.hptb_comm_uses_peer_memory(comm.handle) else "ncclAllGather of accumulators + combine kernel")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_into(Xs, op, axes, out, s=None):
        ax = (c_int32 * len(axes))(*axes)
        st = hb.get_stream() if s is None else s
        if comm is None:
            _ffi.check(hb.lib.hptb_reduce(ctx.handle, _ffi.REDUCE_OPS[op], byref(Xs._c()), ax, len(axes), byref(out._c()), 1, st))
        else:
            _ffi.check(hb.lib.hptb_reduce_sharded(comm.handle, _ffi.REDUCE_OPS[op], byref(Xs._c()), ax, len(axes), 0, row0, ROWS,
                                                   byref(out._c()), st))

    o_sum, o_mean, o_col = T.empty((1,), hb.F32, local), T.empty((1,), hb.F32, local), T.empty((COLS,), hb.F32, local)
    ops = [("sum()", "sum", [0, 1], o_sum, "reduce_rows_kernel"), ("mean()", "mean", [0, 1], o_mean, "reduce_rows_kernel"),
           ("sum(axis 0)", "sum", [0], o_col, "reduce_cols_lean_kernel")]
    op_bytes = [IN_BYTES + 4, IN_BYTES + 4, IN_BYTES + COLS * 4]  # whole job, all ranks

    def step(ev=None, Xs=X):
        for i, (_, op, axes, out, _) in enumerate(ops):
            if ev is not None:
                ev[i].record(stream)
            reduce_into(Xs, op, axes, out)
        if ev is not None:
            ev[len(ops)].record(stream)

    W = max(args.warmup, 3)
    for _ in range(W):
        step()
    barrier()

    # ---- parity guard (full parity lives in tests/): the bench never times wrong results ----------------------------
    # f64 accumulation of the same data with torch on the device, combined over ranks; both readings of north_star's
    # bound are printed: relative to |Σx| (as written) and relative to Σ|x| (the bound that survives cancellation).
    n_total = ROWS * COLS
    ref = torch.stack([big.sum(dtype=torch.float64), big.abs().sum(dtype=torch.float64)])
    ref_col = torch.stack([big.sum(dim=0, dtype=torch.float64), big.abs().sum(dim=0, dtype=torch.float64)])
    if world > 1:
        dist.all_reduce(ref)
        dist.all_reduce(ref_col)
    torch.cuda.synchronize()
    got_sum = float(o_sum.to_cpu().item())
    got_mean = float(o_mean.to_cpu().item())
    got_col = o_col.to_cpu().to(torch.float64)
    rs, ra = ref[0].item(), ref[1].item()
    bound, bound_col = 1e-6 * math.log2(n_total), 1e-6 * math.log2(ROWS)
    col_err = (got_col - ref_col[0].cpu()).abs()
    parity = {
        "sum_rel_to_abs_sum": abs(got_sum - rs) / ra, "sum_rel_to_sum": abs(got_sum - rs) / max(abs(rs), 1e-300),
        "mean_rel_to_abs_sum": abs(got_mean * n_total - rs) / ra,
        "sum0_max_rel_to_abs_sum": float((col_err / ref_col[1].cpu()).max()),
        "sum0_max_rel_to_sum": float((col_err / ref_col[0].cpu().abs().clamp_min(1e-300)).max()),
        "bound_1e-6_log2n": bound, "bound_sum0": bound_col,
    }
    parity["ok"] = bool(parity["sum_rel_to_abs_sum"] <= bound and parity["mean_rel_to_abs_sum"] <= bound + 2 ** -23
                        and parity["sum0_max_rel_to_abs_sum"] <= bound_col)
    ok_t = torch.tensor([1 if parity["ok"] else 0], device=dev)
    if world > 1:
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    parity["ok_all_ranks"] = bool(ok_t.item())
    assert parity["ok_all_ranks"], f"bench parity guard failed: {parity}"
    del ref_col, col_err

    # ---- device-resident timing -------------------------------------------------------------------------------------
    K = args.steps
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(ops) + 1)] for _ in range(K)]
    launches0 = hb.lib.hptb_kernel_launches()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for k in range(K):
        step(evs[k])
    t_end.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = hb.lib.hptb_kernel_launches() - launches0
    ms_total = t_start.elapsed_time(t_end)
    per_op_ms = [sum(evs[k][i].elapsed_time(evs[k][i + 1]) for k in range(K)) / K for i in range(len(ops))]

    # ---- e2e: host buffers in, host buffers out, every step, through the public API (Tensor.to_cuda / reduce / to_cpu)
    shard_bytes = rows_local * COLS * 4
    avail = mem_available_bytes()
    e2e_rows = rows_local
    if avail and shard_bytes * world + (8 << 30) > avail:  # all ranks pin their shard on this host
        e2e_rows = max(1024, int((avail - (8 << 30)) / world / (COLS * 4)) // 1024 * 1024)
    host_x = None
    while host_x is None:
        try:
            host_x = torch.empty((e2e_rows, COLS), dtype=torch.float32, pin_memory=True)
        except RuntimeError:  # the box cannot pin that much: halve the sample
            if e2e_rows <= 1024:
                raise
            e2e_rows = max(1024, e2e_rows // 2 // 1024 * 1024)
    e2e_note = None if e2e_rows == rows_local else f"host memory holds only {e2e_rows} of {rows_local} rows per rank; value scaled to the sample"
    host_x.copy_(big[:e2e_rows])  # synthetic input placed in host memory once, outside the timed region
    torch.cuda.synchronize()
    pinned_out = [torch.empty((1,), dtype=torch.float32, pin_memory=True), torch.empty((1,), dtype=torch.float32, pin_memory=True),
                  torch.empty((COLS,), dtype=torch.float32, pin_memory=True)]
    h2d = host_x.numel() * 4
    d2h = sum(t.numel() * t.element_size() for t in pinned_out)
    e2e_steps = max(2, min(K, 5))
    e2e_full = e2e_rows == rows_local

    def e2e_step():
        Xd = T.to_cuda(host_x, local, sync=False)  # pinned → device, ordered on the compute stream
        if e2e_full:
            for (_, op, axes, out, _), h in zip(ops, pinned_out):
                reduce_into(Xd, op, axes, out)
                out.to_cpu(out=h, sync=False)
        else:  # bounded sample: local reductions of the row block (no exchange: the block is not a shard of the full tensor)
            for (_, op, axes, out, _), h in zip(ops, pinned_out):
                ax = (c_int32 * len(axes))(*axes)
                _ffi.check(hb.lib.hptb_reduce(ctx.handle, _ffi.REDUCE_OPS[op], byref(Xd._c()), ax, len(axes), byref(out._c()), 1, hb.get_stream()))
                out.to_cpu(out=h, sync=False)
        ctx.synchronize()  # the results are in host memory when the step ends
        return Xd

    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        e2e_step()
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    if e2e_full:
        assert abs(float(pinned_out[0].item()) - rs) / ra <= bound, "e2e parity guard: sum mismatch"
    del host_x

    # ---- max over ranks ------------------------------------------------------------------------------------------------
    times = torch.tensor([ms_total, e2e_ms] + per_op_ms, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    times = times.tolist()
    ms_total, e2e_ms, per_op_ms = times[0], times[1], times[2:]
    ms_step = ms_total / K
    value = BYTES_STEP / (ms_step * 1e-3) / 1e9
    e2e_val = BYTES_STEP * (e2e_rows / rows_local) / (e2e_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()

    kernels = []
    for (name, _, _, _, kern), nbytes, ms in zip(ops, op_bytes, per_op_ms):
        gbs = nbytes / (ms * 1e-3) / 1e9  # whole job
        kernels.append({"op": name, "kernel": kern, "us": round(ms * 1e3, 2), "algorithmic_bytes_all_ranks": nbytes,
                        "gbs": round(gbs, 1), "frac_of_peak_x_n": round(gbs / (peak * world), 4), "share_of_step": round(ms / sum(per_op_ms), 4)})
    # dominant kernel: reduce_rows_kernel<sum|mean, f32, VEC 4> (two of the three launches, 2/3 of the step): every rank
    # streams its 17.18/N GB shard once per launch; at N > 1 the launch time includes the exchange in its epilogue
    rows_ms = (per_op_ms[0] + per_op_ms[1]) / 2
    launch_bytes = shard_bytes + 4
    achieved = launch_bytes / (rows_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and world == 1:
        try:
            traffic = json.load(open(tp)).get("reduce_rows_kernel_sum_f32_cfg5_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "reduce_rows_kernel<sum|mean, f32, VEC 4> (split full reduction + fused exchange epilogue)",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": launch_bytes,
                "avg_launch_us": round(rows_ms * 1e3, 2),
                "note": "per rank; a read-only stream can exceed the measured COPY bandwidth (peak counts read+write bytes)"}

    line = {"metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": CONFIG,
            "gelem_per_s": round(ELEMS_STEP / (ms_step * 1e-3) / 1e9, 2),
            "frac_of_measured_peak_x_n": round(value / (peak * world), 4),
            "collective": exchange, "api": "hptb_reduce_sharded" if comm is not None else "hptb_reduce",
            "roofline": roofline, "kernels": kernels, "parity": parity, "clocks": clocks,
            "e2e": {"value": round(e2e_val, 2), "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_ms, 3), "steps": e2e_steps,
                    "pcie_h2d_gbs_per_gpu": round(h2d / (e2e_ms * 1e-3) / 1e9, 2), **({"sample": e2e_note} if e2e_note else {})},
            "gpu_launches": int(launches)}
    rows = None
    if not args.no_rows:
        from bench_rows import run_rows
        rows = run_rows(hb, torch, dist, world, rank, local, stream, peak)
    if rank == 0:
        if world == 1:
            try:
                gbs, dt, cores = time_cpu(12, 2, CPU_SAMPLE_ROWS_BASE)
                line["cpu_baseline"] = {"value": round(gbs, 3), "unit": "GB/s", "cores": cores, "kind": "port",
                                        "sample": f"rows [0,{CPU_SAMPLE_ROWS_BASE}) of the tensor (2.1 GB per pass, the shard of an 8-GPU run), 12 timed "
                                                  "steps of sum(), mean(), sum(axis 0) by the C++/OpenMP restatement of Hpt's CPU path "
                                                  "(oracle/oracle_cpu.cpp); ms/step %.1f" % (dt * 1e3)}
            except Exception as ex:  # the baseline must never hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
        if rows is not None:
            line["rows"] = rows
        print(json.dumps(line), flush=True)
    if comm is not None:
        comm.destroy()
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """The driver parses ONE JSON line from stdout; NCCL and friends print banners there ("NCCL version …").
    Keep a private handle on the real stdout for the JSON line and point fd 1 at stderr for everything else."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    global print
    out = _claim_stdout()
    _print = print

    def print(*a, **k):  # noqa: A001 — the JSON line goes to the real stdout
        k.pop("flush", None)
        _print(*a, file=out, flush=True, **k)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-rows", action="store_true", help="skip the rows for BASELINE configs 1-4")
    ap.add_argument("--rows", action="store_true", help="(default now; kept for old command lines)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
