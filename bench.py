#!/usr/bin/env python
"""bench.py — the measured hot path (BASELINE.json metric) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rows]

Workload (config.workload "cfg2"): BASELINE.json configs[1] — f32 X[8192,8192] contiguous, V = X.t() (a
non-contiguous view, strides [1, 8192]); one step = V.sin(), V.exp(), V.max(0), V.argmax(0) through the
public API (hpt_b200.Tensor → C ABI → sm_100a kernels).  Algorithmic bytes per step:
2·536,870,912 + 268,468,224 + 268,500,992 = 1,610,711,040 B; the input (268 MB) is larger than the 126 MB L2.
`value` = Σ over ranks of algorithmic bytes ÷ max-over-ranks device time (GB/s), inputs resident in HBM.
`e2e`   = same metric with HOST buffers: every step copies X from pinned host memory, runs the four ops and
          copies all four results back to pinned host memory, all inside the timed region.
At N > 1 every rank owns one [8192,8192] row block of a [8192·N, 8192] tensor (outer-axis sharding): the
four ops need no exchange (the reduced view axis is not the sharded one) → "scaling": "weak".  The sharded
full sum of config 5 (NCCL allreduce of partials) is reported beside it under "rows" with --rows.
`--impl reference` times the CPU restatement of Hpt's path (oracle/oracle_cpu.cpp, OpenMP + libmvec) on the
host cores for the same workload; Hpt itself is Rust and cannot be built in this image.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_SIDE = 8192
BYTES_UNARY = 2 * N_SIDE * N_SIDE * 4                  # read + write
BYTES_MAX = N_SIDE * N_SIDE * 4 + N_SIDE * 4
BYTES_ARGMAX = N_SIDE * N_SIDE * 4 + N_SIDE * 8
BYTES_STEP = 2 * BYTES_UNARY + BYTES_MAX + BYTES_ARGMAX  # 1,610,711,040
ELEMS_STEP = 4 * N_SIDE * N_SIDE
METRIC = "hbm_gbs_transposed_unary_axis_reduce"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
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
def load_cpu_port():
    path = os.path.join(ROOT, "oracle", "_build", "liboracle_cpu.so")
    if not os.path.exists(path):
        import build as _b
        _b.build_oracle()
    L = ctypes.CDLL(path)
    vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    L.orc_unary_f32_strided2d.argtypes = [ci, vp, i64, i64, i64, i64, vp]
    L.orc_max_f32_axis0.argtypes = [vp, i64, i64, i64, i64, vp]
    L.orc_argmax_f32_axis0.argtypes = [vp, i64, i64, i64, i64, vp]
    L.orc_sum_f32_all.argtypes = [vp, i64]
    L.orc_sum_f32_all.restype = ctypes.c_float
    L.orc_num_threads.restype = ci
    return L


def cpu_step(L, x, outs, n):
    """the four ops of the workload on the host: logical results of the transposed view."""
    s, e, m, a = outs
    L.orc_unary_f32_strided2d(0, x.ctypes.data, n, n, 1, n, s.ctypes.data)
    L.orc_unary_f32_strided2d(1, x.ctypes.data, n, n, 1, n, e.ctypes.data)
    L.orc_max_f32_axis0(x.ctypes.data, n, n, 1, n, m.ctypes.data)
    L.orc_argmax_f32_axis0(x.ctypes.data, n, n, 1, n, a.ctypes.data)


def time_cpu(steps, warmup, n=N_SIDE):
    import numpy as np
    L = load_cpu_port()
    rng = np.random.default_rng(1234 + 2)
    x = rng.standard_normal((n, n), dtype=np.float32)
    outs = (np.empty((n, n), np.float32), np.empty((n, n), np.float32), np.empty(n, np.float32), np.empty(n, np.int64))
    for _ in range(warmup):
        cpu_step(L, x, outs, n)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(L, x, outs, n)
    dt = (time.perf_counter() - t0) / steps
    scale = (n / N_SIDE) ** 2
    return BYTES_STEP * scale / dt / 1e9, dt, L.orc_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    gbs, dt, cores = time_cpu(max(1, min(args.steps, 5)), max(1, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": METRIC, "value": round(gbs, 3), "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: f32 [8192,8192] transposed view: sin, exp, max(0), argmax(0)",
                       "l2": "input 268 MB > L2"},
            "cpu_baseline": {"value": round(gbs, 3), "unit": "GB/s", "cores": cores, "kind": "port",
                             "sample": f"full workload, {max(1, min(args.steps, 5))} timed steps of the C++/OpenMP restatement "
                                       "(oracle/oracle_cpu.cpp); Hpt's Rust CPU path cannot be built here"},
            "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: hpt_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import hpt_b200 as hb

    stream = torch.cuda.current_stream()
    hb.set_stream(stream.cuda_stream)
    ctx = hb.context(local)
    n = N_SIDE
    # synthetic input, seeded per rank; generated on the host and copied once (outside the timed region)
    g = torch.Generator().manual_seed(1234 + 2 + rank)
    host_x = torch.empty((n, n), dtype=torch.float32).pin_memory()
    host_x.normal_(generator=g)
    X = hb.Tensor.to_cuda(host_x, local)
    V = X.t()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ops = [("sin", lambda: V.sin(), BYTES_UNARY), ("exp", lambda: V.exp(), BYTES_UNARY),
           ("max", lambda: V.max([0]), BYTES_MAX), ("argmax", lambda: V.argmax([0]), BYTES_ARGMAX)]

    def step(ev=None):
        res = []
        for i, (_, fn, _) in enumerate(ops):
            if ev is not None:
                ev[i].record(stream)
            res.append(fn())
        if ev is not None:
            ev[len(ops)].record(stream)
        return res

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # quick parity guard on a slice (full parity lives in tests/): the bench never times wrong results
    r = step()
    torch.cuda.synchronize()
    chk_rows = 64
    sl = host_x[:chk_rows].numpy()  # memory rows = view columns
    got_max = r[2].to_cpu().numpy()[:chk_rows]
    assert (got_max == sl.max(axis=1)).all(), "bench parity guard: max mismatch"
    assert (r[3].to_cpu().numpy()[:chk_rows] == sl.argmax(axis=1)).all(), "bench parity guard: argmax mismatch"
    got_sin = r[0][:, :chk_rows].to_cpu().numpy()
    ref_sin = np.sin(sl.T.astype(np.float64)).astype(np.float32)
    assert np.abs(got_sin.view(np.int32).astype(np.int64) - ref_sin.view(np.int32).astype(np.int64)).max() <= 2
    del r

    K = args.steps
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(ops) + 1)] for _ in range(K)]
    launches0 = hb.lib.hptb_kernel_launches()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for k in range(K):
        step(evs[k])
    t_end.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = hb.lib.hptb_kernel_launches() - launches0
    ms_total = t_start.elapsed_time(t_end)
    per_op_ms = [sum(evs[k][i].elapsed_time(evs[k][i + 1]) for k in range(K)) / K for i in range(len(ops))]

    # e2e: host buffers in, host buffers out, every step — through the public API (Tensor.to_cuda / ops / to_cpu).
    # Three streams (upload, compute, read-back) so that step k's read-back overlaps step k+1's upload and
    # kernels; every byte still crosses PCIe inside the timed region.  Outputs are double-buffered in pinned memory.
    import collections
    pinned_out = [[torch.empty((n, n), dtype=torch.float32).pin_memory(), torch.empty((n, n), dtype=torch.float32).pin_memory(),
                   torch.empty((n,), dtype=torch.float32).pin_memory(), torch.empty((n,), dtype=torch.int64).pin_memory()]
                  for _ in range(2)]
    h2d = host_x.numel() * 4
    d2h = sum(t.numel() * t.element_size() for t in pinned_out[0])
    e2e_steps = max(3, min(K, 10))
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    sc = stream.cuda_stream
    ring = collections.deque()

    def e2e_step(k):
        Xd = hb.Tensor.to_cuda(host_x, local, stream=s_in.cuda_stream, sync=False)
        hb._ffi.check(hb.lib.hptb_stream_wait_stream(ctx.handle, sc, s_in.cuda_stream))
        Vd = Xd.t()
        outs = [Vd.sin(), Vd.exp(), Vd.max([0]), Vd.argmax([0])]
        hb._ffi.check(hb.lib.hptb_stream_wait_stream(ctx.handle, s_out.cuda_stream, sc))
        for o, h in zip(outs, pinned_out[k % 2]):
            o.to_cpu(stream=s_out.cuda_stream, out=h, sync=False)
        ev = torch.cuda.Event()
        ev.record(s_out)
        ring.append((Xd, outs, ev))
        if len(ring) > 1:  # at most two steps in flight: step k-1 must have landed before its buffers are reused
            old = ring.popleft()
            old[2].synchronize()

    def e2e_drain():
        while ring:
            ring.popleft()[2].synchronize()

    e2e_step(0)
    e2e_step(1)
    e2e_drain()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(e2e_steps):
        e2e_step(k)
    e2e_drain()
    stream.wait_stream(s_out)
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    # the read-back results of the last step are checked against the device-resident run's guard values
    assert (pinned_out[(e2e_steps - 1) % 2][2].numpy()[:chk_rows] == sl.max(axis=1)).all(), "e2e parity guard: max mismatch"

    # max over ranks
    times = torch.tensor([ms_total, e2e_ms] + per_op_ms, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    times = times.tolist()
    ms_total, e2e_ms, per_op_ms = times[0], times[1], times[2:]
    ms_step = ms_total / K
    value = world * BYTES_STEP / (ms_step * 1e-3) / 1e9
    e2e_val = world * BYTES_STEP / (e2e_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()

    kernels = []
    for (name, _, nbytes), ms in zip(ops, per_op_ms):
        gbs = nbytes / (ms * 1e-3) / 1e9
        kernels.append({"op": name, "kernel": "map_tiled_smem_kernel" if name in ("sin", "exp") else "reduce_rows_lean_kernel",
                        "us": round(ms * 1e3, 2), "algorithmic_bytes": nbytes, "gbs": round(gbs, 1), "frac": round(gbs / peak, 4),
                        "share_of_step": round(ms / sum(per_op_ms), 4)})
    # dominant kernel: the transposing tile map (sin + exp launches, 60 % of the step)
    tiled_ms = (per_op_ms[0] + per_op_ms[1]) / 2
    achieved = BYTES_UNARY / (tiled_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("map_tiled_smem_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "map_tiled_smem_kernel<sin|exp, f32> (transposed read → shared-memory transpose → contiguous write)",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_UNARY,
                "avg_launch_us": round(tiled_ms * 1e3, 2)}

    line = {"metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "cfg2: f32 [8192,8192] transposed view: sin, exp, max(0), argmax(0)",
                       "shard": "one [8192,8192] row block per GPU (outer-axis sharding), no data-path collective",
                       "l2": "input 268 MB and each output 268 MB exceed the 126 MB L2 (no flush needed)",
                       "gelem_per_s": round(world * ELEMS_STEP / (ms_step * 1e-3) / 1e9, 2)},
            "roofline": roofline, "kernels": kernels, "clocks": clocks,
            "e2e": {"value": round(e2e_val, 2), "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_ms, 3), "steps": e2e_steps},
            "gpu_launches": int(launches)}
    if rank == 0:
        if world == 1:
            try:
                gbs, dt, cores = time_cpu(2, 1)
                line["cpu_baseline"] = {"value": round(gbs, 3), "unit": "GB/s", "cores": cores, "kind": "port",
                                        "sample": "full workload, 2 timed steps of the C++/OpenMP restatement of Hpt's CPU path "
                                                  "(oracle/oracle_cpu.cpp); ms/step %.1f" % (dt * 1e3)}
            except Exception as ex:  # the baseline must never hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
        if args.rows:
            line["rows"] = extra_rows(hb, torch, dist, world, rank, local, stream, peak)
        print(json.dumps(line), flush=True)
    elif args.rows:
        extra_rows(hb, torch, dist, world, rank, local, stream, peak)
    if world > 1:
        dist.destroy_process_group()


def extra_rows(hb, torch, dist, world, rank, local, stream, peak):
    """Other BASELINE configs (device-resident, rotating ≥3 buffer sets where the working set fits L2)."""
    from bench_rows import run_rows
    return run_rows(hb, torch, dist, world, rank, local, stream, peak)


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
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", action="store_true", help="also report the other BASELINE configs (cfg1,3,4,5)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
