"""Throughput of the 2-D flagger on the device (developer tool, GPU only).  The comparison with
the reference's numba implementation on the host cores is `python bench.py --twodflag`.

    python tools/time_twodflag.py [n_time n_freq n_bl] [--reps R]
"""
import argparse
import ctypes
import json
import os
import sys
import time
from ctypes import byref, c_size_t

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from katsdpsigproc_b200 import _capi, accel  # noqa: E402
from katsdpsigproc_b200.rfi import twodflag  # noqa: E402


def make_input(rs, shape):
    n_time, n_freq, n_bl = shape
    data = (4.0 + np.sin(np.linspace(0, 6, n_freq))[None, :, None]
            + rs.standard_normal(shape).astype(np.float32) * 0.1).astype(np.float32)
    data[n_time // 4:n_time // 4 + 2, n_freq // 3:n_freq // 2, :] += 0.7
    data[:, n_freq // 5, :] += 0.5
    data[rs.random_sample(shape) < 0.003] += 3.0
    phase = rs.random_sample(shape).astype(np.float32) * np.float32(2 * np.pi)
    vis = (data * np.exp(1j * phase)).astype(np.complex64)
    flags = rs.random_sample(shape) < 0.02
    return vis, flags


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("shape", type=int, nargs="*", default=[16, 4096, 1024])
    ap.add_argument("--cpu-baselines", type=int, default=0, help="ignored (see bench.py --twodflag)")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    shape = tuple(args.shape)
    rs = np.random.RandomState(1)
    vis, flags = make_input(rs, shape)
    n = vis.size
    context = accel.create_some_context(interactive=False)
    queue = context.create_command_queue()
    flagger = twodflag.SumThresholdFlagger(context=context)
    t0 = time.perf_counter()
    out = flagger.get_flags(vis, flags)
    first = time.perf_counter() - t0
    t0 = time.perf_counter()
    out = flagger.get_flags(vis, flags)
    e2e = time.perf_counter() - t0
    # device time of the kernels alone
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
    phases = (ctypes.c_ulonglong * 20)()
    for rep in range(args.reps + 1):
        if rep == args.reps:
            _capi.call("ksp_twodflag_phases", phases, 0, 1)
        a = queue.enqueue_marker()
        _capi.call("ksp_twodflag", ctypes.c_void_p(queue.stream), byref(p), ctypes.c_void_p(d_vis.buffer.ptr),
                   ctypes.c_void_p(d_fl.buffer.ptr), ctypes.c_void_p(d_out.buffer.ptr),
                   ctypes.c_void_p(d_scr.buffer.ptr), c_size_t(per_bl * batch), ctypes.c_int64(batch))
        b = queue.enqueue_marker()
        queue.finish()
        times.append(b.time_since(a))
    _capi.call("ksp_twodflag_phases", phases, 20, 0)
    names = ["spectrum median", "spectrum gaussians", "spectrum mad", "spectrum interpolate", "spectrum st thresholds",
             "spectrum st lines", "2d gaussians", "2d mad", "2d interpolate", "st time", "st freq thresholds",
             "st freq lines", "combine", "unaverage + fill", "-", "elementwise", "box time passes", "box freq passes",
             "gaussian mask + normalise", "-"]
    tot = float(sum(phases)) or 1.0
    phase_tab = {n: round(int(c) / tot, 4) for n, c in zip(names, phases) if n != "-"}
    phase_tab["cycles of block 0"] = int(tot)
    dev = sorted(times[1:])[len(times[1:]) // 2]
    res = {"shape": list(shape), "samples": n, "flagged_fraction": float(out.mean()),
           "gpu_device_s": dev, "gpu_device_Msamples_s": n / dev / 1e6,
           "gpu_end_to_end_s": e2e, "gpu_end_to_end_Msamples_s": n / e2e / 1e6, "first_call_s": first,
           "scratch_MB_per_baseline": per_bl / 1e6, "baselines_in_flight": batch, "phase_share_block0_last_rep": phase_tab}
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.environ.get("TK_OUT", "gpurun_out/time_twodflag.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
