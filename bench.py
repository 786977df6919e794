#!/usr/bin/env python
"""Benchmark of the RFI-flagging hot path (BASELINE.json metric: visibilities flagged/s).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this implementation
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the CPU reference arm

One "step" flags one synthetic dump: at N=1 the BASELINE.json configs[1] workload
(32768 channels x 8320 baselines complex64, median width 13, SumThreshold windows
1..64, n_sigma 11).  With N ranks (torchrun, one process per GPU) every rank flags
its own 8320-baseline range of a 8320*N-baseline dump: baselines are independent,
so there is no data-path collective and scaling is weak.

The JSON line carries: ``value`` (device-resident input, CUDA-event timed, max over
ranks), ``e2e`` (through the public FlaggerDevice API from pinned host buffers,
H2D + kernels + D2H inside the timed region), ``roofline`` (dominant kernel, timed
live with CUDA events on the flagger's stream), ``cpu_baseline`` (the numpy/pandas
port of the reference's FlaggerHost on this box's host cores; rank 0, N=1 only),
``clocks`` and ``gpu_launches``.
"""

from __future__ import annotations

import argparse
import json
import multiprocessing
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "visibilities_flagged_per_second"
UNIT = "vis/s"
CHANNELS = 32768
BASELINES = 8320
WIDTH = 13
N_WINDOWS = 7
N_SIGMA = 11.0
FALLOFF = 1.2
CPU_SAMPLE_BASELINES_PER_CORE = 32


# ----------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """Flag ``baselines`` baselines of a synthetic dump with the numpy/pandas restatement of
    the reference's FlaggerHost (oracle/host_numpy.py); returns (seconds, visibilities)."""
    import warnings

    seed, channels, baselines, start_at = args
    warnings.filterwarnings("ignore")
    from oracle import host_numpy

    vis, _ = host_numpy.synthetic_vis(channels, baselines, seed=seed)
    while time.time() < start_at:       # line the workers up so that they contend as in a real run
        time.sleep(0.001)
    t0 = time.perf_counter()
    host_numpy.flagger(vis, None, width=WIDTH, n_sigma=N_SIGMA, n_windows=N_WINDOWS,
                       threshold_falloff=FALLOFF)
    return time.perf_counter() - t0, channels * baselines


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_flagger_rate(channels: int, steps: int, warmup: int, cores: int):
    """Whole-box throughput of the CPU port: ``cores`` processes, each flagging its own
    baseline range (the same sharding the GPUs use).  Returns (vis/s, ms per step, sample)."""
    per_core = CPU_SAMPLE_BASELINES_PER_CORE
    ctx = multiprocessing.get_context("fork")
    rates, times = [], []
    with ctx.Pool(cores) as pool:
        for step in range(warmup + steps):
            start_at = time.time() + 1.5 + 0.01 * cores
            jobs = [(1000 * step + i + 1, channels, per_core, start_at) for i in range(cores)]
            out = pool.map(_cpu_worker, jobs, chunksize=1)
            wall = max(t for t, _ in out)
            nvis = sum(n for _, n in out)
            if step >= warmup:
                rates.append(nvis / wall)
                times.append(wall)
    sample = (f"{channels} channels x {per_core * cores} baselines per step "
              f"({per_core} per process x {cores} processes), {steps} steps")
    return sum(rates) / len(rates), 1e3 * sum(times) / len(times), sample


def reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    steps = max(1, args.steps)
    value, ms, sample = cpu_flagger_rate(args.channels, steps, args.warmup, cores)
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_ranks: int) -> dict:
    return {
        "workload": (f"FlaggerDevice {args.channels} channels x {args.baselines} baselines "
                     f"complex64 per GPU (BASELINE.json configs[1]), median width {WIDTH}, "
                     f"sum-threshold windows 1..{2 ** (N_WINDOWS - 1)}, n_sigma {N_SIGMA}"),
        "channels": args.channels,
        "baselines_per_gpu": args.baselines,
        "baselines_total": args.baselines * n_ranks,
        "sharding": f"baseline ranges, {n_ranks} rank(s), no collective",
        "l2_policy": "inputs (2.18 GB per dump) exceed the 126 MB L2; no explicit flush",
    }


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU with NVML while a region runs."""

    def __init__(self, index: int, period: float = 0.02) -> None:
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None
        self._period = period

    _BITS = {
        0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def _run(self) -> None:
        nv = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle)
                for bit, name in self._BITS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self._period)

    def __enter__(self):
        if self._nvml is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self) -> dict:
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------- GPU arm
def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def b200_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process: the workers are forked
        cores = host_cores()
        value, ms, sample = cpu_flagger_rate(args.channels, 2, 1, cores)
        cpu = {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
               "what": "oracle/host_numpy.py (numpy/pandas restatement of the reference's "
                       "FlaggerHost), one process per host core"}

    import numpy as np
    import torch
    import torch.distributed as dist

    from katsdpsigproc_b200 import _capi, accel, cuda, sharding, streaming
    from katsdpsigproc_b200.rfi import device as rfi_device

    torch.cuda.set_device(local_rank)
    # one process per GPU: run (and allocate the pinned staging buffers) next to that GPU
    bound_cpus = sharding.bind_to_device_locality(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    C, B = args.channels, args.baselines
    context = cuda.Device(local_rank).make_context()
    queue = context.create_command_queue()
    template = rfi_device.FlaggerDeviceTemplate(
        rfi_device.BackgroundMedianFilterDeviceTemplate(context, WIDTH),
        rfi_device.NoiseEstMADTDeviceTemplate(context, max(C, 10240)),
        rfi_device.ThresholdSumDeviceTemplate(context, n_windows=N_WINDOWS),
        fused=not args.unfused)
    flagger = template.instantiate(queue, C, B, threshold_args={"n_sigma": N_SIGMA,
                                                                "threshold_falloff": FALLOFF})
    flagger.ensure_all_bound()
    vis_dev = flagger.buffer("vis")
    flags_dev = flagger.buffer("flags")

    # synthetic dump of this rank's baseline range, generated on the device (torch = plumbing)
    torch.manual_seed(1 + rank)
    stride = vis_dev.padded_shape[1]
    gen = torch.zeros(C, stride, 2, device="cuda", dtype=torch.float32)
    gen[:, :B].normal_()
    hit = torch.rand(C, B, device="cuda") < (1.0 / 64.0)
    hit |= torch.rand(C, 1, device="cuda") < 0.005
    amp = torch.rand(C, B, device="cuda") * 20.0 + 50.0
    phase = torch.rand(C, B, device="cuda") * (2.0 * np.pi)
    gen[:, :B, 0] += hit * amp * torch.cos(phase)
    gen[:, :B, 1] += hit * amp * torch.sin(phase)
    injected = float(hit.float().mean())
    del hit, amp, phase
    torch.cuda.synchronize()
    _capi.call("ksp_memcpy_async", vis_dev.ptr, gen.data_ptr(), gen.numel() * 4, _capi.D2D,
               queue.stream)
    queue.finish()

    # end-to-end leg: the streaming front end (pinned host staging, upload / compute /
    # download on three queues, two dumps in flight); both staging buffers hold the dump
    stream = streaming.StreamingFlagger(template, C, B, depth=2,
                                        threshold_args={"n_sigma": N_SIGMA,
                                                        "threshold_falloff": FALLOFF})
    first = vis_dev.get(queue, stream.host_vis(0))
    for k in range(1, stream.depth):
        np.copyto(stream.host_vis(k), first)
    del gen
    torch.cuda.empty_cache()

    n_vis = C * B
    sampler = ClockSampler(local_rank)

    def reduce_max(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def timed(step_fn, steps):
        barrier()
        start = queue.enqueue_marker()
        for _ in range(steps):
            step_fn()
        stop = queue.enqueue_marker()
        queue.finish()
        barrier()
        return reduce_max(1e3 * stop.time_since(start) / steps)

    # ---- device-resident throughput
    for _ in range(args.warmup):
        flagger()
    queue.finish()
    launches0 = _capi.kernel_launch_count()
    with sampler:
        ms_step = timed(flagger, args.steps)
        launches = _capi.kernel_launch_count() - launches0

        # ---- end to end through the public API: every step uploads the dump from pinned host
        # memory, flags it and downloads the flags; steps overlap inside StreamingFlagger
        for _ in range(stream.depth):
            stream.submit(None)
        stream.drain()
        barrier()
        t0 = time.perf_counter()
        flags_host = None
        for _ in range(args.steps):
            out = stream.submit(None)
            flags_host = out if out is not None else flags_host
        for out in stream.drain():
            flags_host = out
        torch.cuda.synchronize()
        ms_e2e = reduce_max(1e3 * (time.perf_counter() - t0) / args.steps)
        barrier()
    flagged = float(np.count_nonzero(flags_host[:, :B])) / n_vis

    # ---- dominant kernel, timed live with CUDA events around every stage launch
    roofline = None
    if template.fused:
        _capi.profile_enable(True)
        for _ in range(args.steps):
            flagger()
        queue.finish()
        stages = _capi.profile_read()
        _capi.profile_enable(False)
        per_unit = {"background": 12.0, "noise": 4.0, "threshold": 4.125, "expand_flags": 1.125}
        top = max(stages, key=lambda k: stages[k][0])
        top_ms, top_launches = stages[top]
        peak, peak_source = measured_peak()
        achieved = per_unit[top] * n_vis * args.steps / (top_ms * 1e-3) / 1e9
        kernel_names = {"background": "bg13_kernel", "noise": "madnz_stream_kernel",
                        "threshold": "threshold_sum_kernel", "expand_flags": "expand_flags_kernel"}
        roofline = {
            "bound": "hbm", "kernel": kernel_names[top],
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": peak_source,
            "traffic": ncu_traffic(kernel_names[top]),
            "algorithmic_bytes_per_vis": per_unit[top],
            "launch_ms": top_ms / max(top_launches, 1),
            "launches_per_step": top_launches / args.steps,
            "stage_ms_per_step": {k: v[0] / args.steps for k, v in stages.items()},
            "pipeline_bytes_per_vis": 9.0,
            "pipeline_achieved": 9.0 * n_vis / (ms_step * 1e-3) / 1e9,
            "pipeline_frac": 9.0 * n_vis / (ms_step * 1e-3) / 1e9 / peak,
        }

    if rank == 0:
        total_vis = n_vis * world
        h2d = int(np.prod(vis_dev.padded_shape)) * vis_dev.dtype.itemsize
        d2h = int(np.prod(flags_dev.padded_shape))
        line = {
            "metric": METRIC, "value": total_vis / (ms_step * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, world), fused=bool(template.fused),
                           injected_fraction=injected, flagged_fraction=flagged),
            "e2e": {"value": total_vis / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e,
                    "how": "StreamingFlagger: pinned host -> device, flagger, flags -> pinned host; "
                           "2 dumps in flight on upload/compute/download queues; wall clock",
                    "cpu_binding": (f"rank 0 bound to {len(bound_cpus)} CPUs next to its GPU (NVML affinity)"
                                    if bound_cpus else "none")},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--channels", type=int, default=CHANNELS)
    ap.add_argument("--baselines", type=int, default=BASELINES)
    ap.add_argument("--unfused", action="store_true",
                    help="run the reference's 5-operation sequence instead of the fused flagger")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
