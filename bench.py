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
H2D + kernels + D2H inside the timed region), ``parity`` (flags and noise of a
deterministic subset of this very dump against the host classes, in the same run),
``roofline`` (the flagger's kernel, timed live with CUDA events on its stream; the
four stages of the chunked form beside it), ``cfg5`` (BASELINE.json configs[4]: ONE
32768 x 12960 dump sharded over the ranks by baseline ranges, blocks of 16 dumps),
``cpu_baseline`` (the reference's FlaggerHost on this box's host cores; rank 0, N=1
only), ``clocks`` and ``gpu_launches``.
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

    run = host_flagger()[0]
    vis, _ = host_numpy.synthetic_vis(channels, baselines, seed=seed)
    while time.time() < start_at:       # line the workers up so that they contend as in a real run
        time.sleep(0.001)
    t0 = time.perf_counter()
    run(vis)
    return time.perf_counter() - t0, channels * baselines


def host_flagger():
    """(callable vis -> flags, kind, description): the UNMODIFIED reference FlaggerHost from
    oracle/_ref when it is there (placed by oracle/make_ref.py where the reference checkout
    exists), else its numpy/pandas restatement oracle/host_numpy.py."""
    import oracle
    from oracle import host_numpy

    ref = oracle.reference_host()
    if ref is not None:
        fn = ref.FlaggerHost(ref.BackgroundMedianFilterHost(WIDTH), ref.NoiseEstMADHost(),
                             ref.ThresholdSumHost(N_SIGMA, n_windows=N_WINDOWS,
                                                  threshold_falloff=FALLOFF))
        return fn, "reference", ("katsdpsigproc.rfi.host.FlaggerHost, unmodified (oracle/_ref), one "
                                 "process per host core")

    def port(vis):
        return host_numpy.flagger(vis, None, width=WIDTH, n_sigma=N_SIGMA, n_windows=N_WINDOWS,
                                  threshold_falloff=FALLOFF)
    return port, "port", ("oracle/host_numpy.py (numpy/pandas restatement of the reference's "
                          "FlaggerHost), one process per host core")


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
    _, kind, what = host_flagger()
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": sample, "what": what},
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


def source_sha16() -> str:
    """Fingerprint of the CUDA sources the FLAGGER's kernels are built from (the helper
    operations and the 2-D flagger are separate translation units and do not enter)."""
    import hashlib

    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "katsdpsigproc_b200", "csrc")
    names = ["Makefile", "common.cuh", "median13.cuh", "select.cuh", "tma.cuh", "bg13.cuh", "background.cu",
             "madnz_stream.cuh", "madnz.cu", "threshold_tile.cuh", "threshold.cu", "flagger.cu",
             "dataflow.h", "dataflow.cu"]
    for name in names:
        with open(os.path.join(csrc, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel: str):
    """(DRAM bytes per launch of `kernel`, note) from the committed ncu capture
    profiles/roofline_traffic.json - only if that capture was taken from THESE sources."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            rec = json.load(f)
    except Exception:
        return None, "no capture committed"
    sha = source_sha16()
    if rec.get("source_sha16") != sha:
        return None, (f"capture {rec.get('tag')} is of sources {rec.get('source_sha16')}, "
                      f"this build is {sha}: not reported")
    return rec.get("kernels", {}).get(kernel), f"ncu --set full capture {rec.get('tag')} of these sources"


def b200_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process: the workers are forked
        cores = host_cores()
        value, ms, sample = cpu_flagger_rate(args.channels, 2, 1, cores)
        _, kind, what = host_flagger()
        cpu = {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
               "what": what}

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

    peak, peak_source = measured_peak()
    stats = dict(flagger.stats()) if template.fused else {}
    dataflow = bool(flagger.parameters().get("dataflow"))

    # ---- parity of THIS dump, in this run: a deterministic subset of baselines at full channel
    # count ({0..63} U every 65th U last 64) against the host classes
    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_block(flagger, queue, C, B)

    # ---- the chunked four-kernel form of the same flagger: the four stages timed one after
    # the other (one lane, CUDA events around every launch) and as it runs (4 lanes)
    chunked = None
    if template.fused and not args.no_stages:
        chunked = chunked_form(template, queue, flagger, C, B, min(args.steps, 5), timed, peak)

    roofline = None
    if template.fused:
        pipe_achieved = 9.0 * n_vis / (ms_step * 1e-3) / 1e9
        if dataflow or not chunked:
            # one kernel does everything: the kernel's roofline is the pipeline's
            kernel = "dataflow_kernel" if dataflow else "flagger"
            per_vis, launch_ms, per_step = 9.0, ms_step / max(launches / args.steps, 1), launches / args.steps
            achieved = pipe_achieved
            timing = "CUDA events around the timed steps"
        else:
            # the dominant kernel of the chunked form, timed live with CUDA events around its launches
            top = max(chunked["stage_ms_per_step"], key=chunked["stage_ms_per_step"].get)
            kernel = chunked["stage_kernel"][top].split(" ")[0]
            per_vis = chunked["stage_algorithmic_bytes_per_vis"][top]
            per_step = chunked["stage_launches_per_step"][top]
            launch_ms = chunked["stage_ms_per_step"][top] / max(per_step, 1)
            achieved = per_vis * n_vis / (chunked["stage_ms_per_step"][top] * 1e-3) / 1e9
            timing = ("CUDA events recorded inside ksp_flagger around every launch of this stage, stages "
                      "one after the other (one lane); the timed steps run 4 lanes, see chunked_form")
        traffic, traffic_note = ncu_traffic(kernel)
        roofline = {
            "bound": "hbm", "kernel": kernel,
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": peak_source,
            "traffic": traffic, "traffic_source": traffic_note,
            "algorithmic_bytes_per_vis": per_vis,
            "launch_ms": launch_ms, "launches_per_step": per_step, "timing": timing,
            "pipeline_bytes_per_vis": 9.0, "pipeline_achieved": pipe_achieved,
            "pipeline_frac": pipe_achieved / peak,
        }
        if dataflow and stats:
            busy = sum(stats[k] for k in ("cycles_background", "cycles_noise", "cycles_threshold",
                                          "cycles_expand")) or 1
            roofline["work_item_share"] = {
                "background": stats["cycles_background"] / busy, "noise": stats["cycles_noise"] / busy,
                "threshold": stats["cycles_threshold"] / busy, "expand_flags": stats["cycles_expand"] / busy,
                "of_which_waiting_for_other_items": stats["cycles_wait"] / busy,
                "what": "block-cycles per kind of work item of the last launch (ksp_flagger_stats)"}
            roofline["noise_rows_redone_by_radix_select"] = stats["fallbacks"]
        if chunked:
            roofline["chunked_form"] = chunked

    # ---- BASELINE.json configs[4]: ONE 32768 x 12960 dump sharded by baseline ranges
    cfg5 = None
    if not args.no_cfg5:
        del stream
        cfg5 = cfg5_block(template, queue, context, rank, world, barrier, reduce_max, args)

    twod = None
    if rank == 0 and args.twodflag:
        twod = twodflag_block(context)

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
                           dataflow=dataflow, injected_fraction=injected, flagged_fraction=flagged),
            "e2e": {"value": total_vis / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e,
                    "how": "StreamingFlagger: pinned host -> device, flagger, flags -> pinned host; "
                           "2 dumps in flight on upload/compute/download queues; wall clock",
                    "cpu_binding": (f"rank 0 bound to {len(bound_cpus)} CPUs next to its GPU (NVML affinity)"
                                    if bound_cpus else "none")},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "parity": parity,
            "roofline": roofline,
            "cfg5": cfg5,
            "twodflag": twod,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def parity_subset(baselines: int):
    import numpy as np

    idx = set(range(min(64, baselines))) | set(range(0, baselines, 65)) | \
        set(range(max(0, baselines - 64), baselines))
    return np.array(sorted(idx))


def parity_block(flagger, queue, C: int, B: int) -> dict:
    """Flags and noise of the benchmarked dump as the device left them, for a subset of
    baselines, against (a) the float32 device contract (oracle/contract.c; must be identical),
    (b) the host classes in float64 - the unmodified reference FlaggerHost when oracle/_ref
    holds it, and its numpy port, which also counts the window decisions that lie within 1e-6
    (relative) of their threshold, the only places where a flag may legitimately differ."""
    import warnings

    import numpy as np

    from oracle import contract
    from oracle import host_numpy as hn

    t0 = time.perf_counter()
    flagger()                                           # the device buffers hold this dump's results
    pick = parity_subset(B)
    vis = np.ascontiguousarray(np.array(flagger.buffer("vis").get(queue))[:, pick])
    flags = np.ascontiguousarray(np.array(flagger.buffer("flags").get(queue))[:, pick])
    noise = np.array(flagger.buffer("noise").get(queue))[pick]
    abs_mode = contract.detect_abs_mode()
    c_flags, _, c_noise = contract.flagger(vis, None, width=WIDTH, n_sigma=N_SIGMA,
                                           n_windows=N_WINDOWS, threshold_falloff=FALLOFF,
                                           abs_mode=abs_mode)
    near, stages = {}, {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        h_flags = hn.flagger(vis, None, width=WIDTH, n_sigma=N_SIGMA, n_windows=N_WINDOWS,
                             threshold_falloff=FALLOFF, stages=stages, near=near)
        run, kind, _ = host_flagger()
        r_flags = run(vis) if kind == "reference" else None
    h_noise = stages["noise"].astype(np.float32)
    ulp = np.abs(noise.view(np.int32).astype(np.int64) - h_noise.view(np.int32).astype(np.int64))
    out = {
        "shape": [C, B], "baselines_checked": int(pick.size),
        "flag_mismatches": int(np.count_nonzero(flags != h_flags)),
        "near_threshold_1e-6": int(near.get("band", 0)),
        "noise_max_ulp": int(ulp.max()),
        "contract_flag_mismatches": int(np.count_nonzero(flags != c_flags)),
        "contract_noise_mismatches": int(np.count_nonzero(noise.view(np.uint32) != c_noise.view(np.uint32))),
        "flags_checked": int(flags.size), "flags_set": int(np.count_nonzero(flags)),
        "host": "oracle/host_numpy.py (FlaggerHost restated, float64)",
        "seconds": None,
    }
    if r_flags is not None:
        out["reference_flag_mismatches"] = int(np.count_nonzero(flags != r_flags))
        out["reference"] = "katsdpsigproc.rfi.host.FlaggerHost, unmodified (oracle/_ref)"
    out["seconds"] = round(time.perf_counter() - t0, 2)
    return out


def chunked_form(template, queue, flagger, C: int, B: int, steps: int, timed, peak: float) -> dict:
    """The same flagger as four launches per chunk of baselines (ksp_flagger with
    chunk_baselines > 0), sharing the main flagger's vis / noise / flags buffers."""
    from ctypes import byref

    from katsdpsigproc_b200 import _capi
    from katsdpsigproc_b200.rfi import device as rfi_device

    probe = rfi_device.FusedFlaggerDevice(template.background, template.threshold, queue, C, B,
                                          N_SIGMA, FALLOFF, chunk_baselines=16 * 148)
    probe.bind(vis=flagger.buffer("vis"), noise=flagger.buffer("noise"),
               out_flags=flagger.buffer("flags"))
    probe.ensure_all_bound()
    for _ in range(2):
        probe()
    queue.finish()
    ms_lanes = timed(probe, steps)
    # stage by stage: one step at a time, the median over the steps of each stage's time (an
    # event pair around every launch also catches whatever delays the host adds in between;
    # the median keeps such a step from colouring the figure)
    _capi.profile_enable(True)
    per_step = []
    for _ in range(max(steps, 7)):
        probe()
        queue.finish()
        per_step.append(_capi.profile_read())
    _capi.profile_enable(False)
    steps = 1
    stages = {}
    for name in per_step[0]:
        times = sorted(rec[name][0] for rec in per_step)
        stages[name] = (times[len(times) // 2], per_step[0][name][1])
    per_unit = {"background": 12.0, "noise": 4.0, "threshold": 4.125, "expand_flags": 1.125}
    names = {"background": "bg13_kernel", "noise": "madnz_stream_kernel",
             "threshold": "threshold_sum_kernel (two passes)", "expand_flags": "expand_flags_kernel"}
    n_vis = C * B
    out = {
        "what": "background filter of the whole dump in one launch, then noise, two threshold passes and flag expansion per chunk of 2368 baselines on 4 lanes; deviations go through device memory",
        "ms_per_step_4_lanes": ms_lanes,
        "ms_per_step_1_lane_sum_of_stages": sum(v[0] for v in stages.values()) / steps,
        "stage_ms_per_step": {k: v[0] / steps for k, v in stages.items()},
        "stage_launches_per_step": {k: v[1] / steps for k, v in stages.items()},
        "stage_kernel": names,
        "stage_algorithmic_bytes_per_vis": per_unit,
        "stage_frac_of_peak": {k: per_unit[k] * n_vis * steps / (v[0] * 1e-3) / 1e9 / peak
                               for k, v in stages.items() if v[0] > 0},
    }
    del probe
    return out


def twodflag_block(context) -> dict:
    """`--twodflag`: the 2-D flagger (rfi.twodflag.SumThresholdFlagger) on a 16 x 4096 x 1024
    complex64 block: device time, time through get_flags with host arrays, and the reference's numba
    implementation (oracle/_ref) with a thread pool on the host cores on 64 of the baselines, flags
    compared."""
    import concurrent.futures
    import ctypes
    from ctypes import byref, c_size_t

    import numpy as np

    import oracle
    from katsdpsigproc_b200 import _capi, accel
    from katsdpsigproc_b200.rfi import twodflag

    shape = (16, 4096, 1024)
    rs = np.random.RandomState(1)
    mag = (4.0 + np.sin(np.linspace(0, 6, shape[1]))[None, :, None]
           + rs.standard_normal(shape).astype(np.float32) * 0.1).astype(np.float32)
    mag[4:6, shape[1] // 3:shape[1] // 2, :] += 0.7
    mag[:, shape[1] // 5, :] += 0.5
    mag[rs.random_sample(shape) < 0.003] += 3.0
    phase = rs.random_sample(shape).astype(np.float32) * np.float32(2 * np.pi)
    vis = (mag * np.exp(1j * phase)).astype(np.complex64)
    flags = rs.random_sample(shape) < 0.02
    flagger = twodflag.SumThresholdFlagger(context=context)
    flagger.get_flags(vis, flags)
    t0 = time.perf_counter()
    out = flagger.get_flags(vis, flags)
    e2e = time.perf_counter() - t0
    queue = context.create_command_queue()
    p = flagger._params(shape, True)
    lib = _capi.load()
    per_bl = int(lib.ksp_twodflag_scratch_bytes(byref(p), 1))
    batch = max(1, min(shape[2], int(lib.ksp_twodflag_resident_baselines()), (8 << 30) // per_bl))
    d_vis = accel.DeviceArray(context, vis.shape, vis.dtype)
    d_fl = accel.DeviceArray(context, vis.shape, np.uint8)
    d_out = accel.DeviceArray(context, vis.shape, np.uint8)
    d_scr = accel.DeviceArray(context, (per_bl * batch,), np.uint8)
    d_vis.set(queue, vis)
    d_fl.set(queue, flags.astype(np.uint8))
    times = []
    for _ in range(4):
        a = queue.enqueue_marker()
        _capi.call("ksp_twodflag", ctypes.c_void_p(queue.stream), byref(p), ctypes.c_void_p(d_vis.buffer.ptr),
                   ctypes.c_void_p(d_fl.buffer.ptr), ctypes.c_void_p(d_out.buffer.ptr),
                   ctypes.c_void_p(d_scr.buffer.ptr), c_size_t(per_bl * batch), ctypes.c_int64(batch))
        m = queue.enqueue_marker()
        queue.finish()
        times.append(m.time_since(a))
    dev = sorted(times[1:])[1]
    res = {"workload": "SumThresholdFlagger.get_flags, 16 dumps x 4096 channels x 1024 baselines complex64, default parameters",
           "samples": vis.size, "flagged_fraction": float(out.mean()),
           "device_s": dev, "device_samples_per_s": vis.size / dev,
           "e2e_s": e2e, "e2e_samples_per_s": vis.size / e2e}
    ref = oracle.reference_twodflag()
    if ref is not None:
        nb = 64
        sub_v, sub_f = np.ascontiguousarray(vis[..., :nb]), np.ascontiguousarray(flags[..., :nb])
        cpu = ref.SumThresholdFlagger()
        cpu.get_flags(sub_v[..., :2], sub_f[..., :2])             # numba compilation
        cores = host_cores()
        with concurrent.futures.ThreadPoolExecutor(cores) as pool:
            t0 = time.perf_counter()
            want = cpu.get_flags(sub_v, sub_f, pool=pool)
            cpu_s = time.perf_counter() - t0
        res["cpu_reference"] = {"kind": "reference", "what": "katsdpsigproc.rfi.twodflag.SumThresholdFlagger (numba), "
                                "unmodified (oracle/_ref), ThreadPoolExecutor", "threads": cores, "baselines": nb,
                                "samples_per_s": sub_v.size / cpu_s}
        res["flag_mismatches_vs_reference"] = int(np.count_nonzero(want != out[..., :nb]))
        res["baselines_compared"] = nb
    return res


CFG5_BASELINES = 12960
CFG5_DUMPS = 16


def cfg5_block(template, queue, context, rank: int, world: int, barrier, reduce_max, args) -> dict:
    """BASELINE.json configs[4]: one 32768 x 12960 dump (80 antennas, 4 polarisations) split
    into contiguous baseline ranges, one per rank (sharding.baseline_ranges), flagged in blocks
    of 16 dumps.  Device-resident time per dump (max over ranks), end to end through
    StreamingFlagger, and - with several ranks - the whole dump on rank 0 alone, timed in the
    same run, for the strong-scaling ratio."""
    import numpy as np
    import torch

    from katsdpsigproc_b200 import _capi, sharding, streaming

    C = args.channels
    start, stop = sharding.baseline_ranges(CFG5_BASELINES, world)[rank]
    nb = stop - start
    targs = {"n_sigma": N_SIGMA, "threshold_falloff": FALLOFF}

    def make(nbl, seed):
        fl = template.instantiate(queue, C, nbl, threshold_args=targs)
        fl.ensure_all_bound()
        vis = fl.buffer("vis")
        stride = vis.padded_shape[1]
        torch.manual_seed(seed)
        gen = torch.zeros(C, stride, 2, device="cuda", dtype=torch.float32)
        gen[:, :nbl].normal_()
        hit = torch.rand(C, nbl, device="cuda") < (1.0 / 64.0)
        gen[:, :nbl, 0] += hit * (torch.rand(C, nbl, device="cuda") * 20.0 + 50.0)
        del hit
        torch.cuda.synchronize()
        _capi.call("ksp_memcpy_async", vis.ptr, gen.data_ptr(), gen.numel() * 4, _capi.D2D, queue.stream)
        queue.finish()
        del gen
        torch.cuda.empty_cache()
        return fl

    def time_block(fl, dumps):
        for _ in range(3):
            fl()
        queue.finish()
        barrier()
        a = queue.enqueue_marker()
        for _ in range(dumps):
            fl()
        b = queue.enqueue_marker()
        queue.finish()
        barrier()
        return 1e3 * b.time_since(a) / dumps

    fl = make(nb, 100 + rank)
    ms_shard = reduce_max(time_block(fl, CFG5_DUMPS))
    out = {
        "workload": (f"ONE {C} x {CFG5_BASELINES} complex64 dump sharded by baseline ranges over "
                     f"{world} rank(s), blocks of {CFG5_DUMPS} dumps (BASELINE.json configs[4])"),
        "baselines_this_rank": nb, "shards": sharding.shard_sizes(CFG5_BASELINES, world),
        "dumps_per_block": CFG5_DUMPS,
        "ms_per_dump": ms_shard,
        "value": C * CFG5_BASELINES / (ms_shard * 1e-3), "unit": UNIT,
        "dataflow": bool(fl.parameters().get("dataflow")),
        "launches_per_dump": None,
    }
    n0 = _capi.kernel_launch_count()
    fl()
    queue.finish()
    out["launches_per_dump"] = _capi.kernel_launch_count() - n0
    # end to end: every dump of the block uploaded from pinned host memory, flags downloaded
    stream = streaming.StreamingFlagger(template, C, nb, depth=2, threshold_args=targs)
    first = fl.buffer("vis").get(queue, stream.host_vis(0))
    for k in range(1, stream.depth):
        np.copyto(stream.host_vis(k), first)
    del fl
    for _ in range(stream.depth):
        stream.submit(None)
    stream.drain()
    barrier()
    t0 = time.perf_counter()
    for _ in range(CFG5_DUMPS):
        stream.submit(None)
    stream.drain()
    torch.cuda.synchronize()
    ms_e2e = reduce_max(1e3 * (time.perf_counter() - t0) / CFG5_DUMPS)
    barrier()
    out["e2e"] = {"ms_per_dump": ms_e2e, "value": C * CFG5_BASELINES / (ms_e2e * 1e-3), "unit": UNIT,
                  "h2d_bytes_per_dump_this_rank": int(np.prod(first.shape)) * 8}
    del stream, first
    if world > 1:
        # the whole dump on ONE GPU, same run: rank 0 works, the others wait at the barrier
        ms_full = 0.0
        if rank == 0:
            full = make(CFG5_BASELINES, 100)
            for _ in range(3):
                full()
            queue.finish()
            a = queue.enqueue_marker()
            for _ in range(CFG5_DUMPS):
                full()
            b = queue.enqueue_marker()
            queue.finish()
            ms_full = 1e3 * b.time_since(a) / CFG5_DUMPS
            del full
        barrier()
        ms_full = reduce_max(ms_full)
        out["ms_per_dump_whole_dump_on_one_gpu"] = ms_full
        out["strong_scaling_speedup"] = ms_full / ms_shard
        out["strong_scaling_efficiency"] = ms_full / ms_shard / world
    return out


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
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity block")
    ap.add_argument("--no-stages", action="store_true",
                    help="skip the chunked form's per-stage timings")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the configs[4] block")
    ap.add_argument("--twodflag", action="store_true",
                    help="also measure the 2-D flagger against the reference's numba implementation")
    ap.add_argument("--quick", action="store_true",
                    help="only the two timed legs: no CPU baseline, parity, stages or cfg5")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)
    if args.quick:
        args.no_cpu_baseline = args.no_parity = args.no_stages = args.no_cfg5 = True
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
