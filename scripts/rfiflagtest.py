#!/usr/bin/env python
"""1-D RFI flagging benchmark on random or real data.

Command-line twin of the reference's ``scripts/rfiflagtest.py`` (1-D part, lines 47-108 and
135-207) running on this package: same presets, same options, same printed timings, so the
two can be run side by side.  The reference's ``--host`` and ``--times`` (CPU / 2-D flagger)
paths are not part of this package.
"""

import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from katsdpsigproc_b200 import accel  # noqa: E402
from katsdpsigproc_b200.rfi import device as rfi_device  # noqa: E402


def generate_data(channels, baselines):
    """Unit-variance complex normal noise, one channel row at a time from RandomState(1)."""
    rs = np.random.RandomState(seed=1)
    out = np.empty((channels, baselines), np.complex64)
    for row in out:
        real = rs.standard_normal(size=baselines).astype(np.float32)
        imag = rs.standard_normal(size=baselines).astype(np.float32)
        row[:] = real + 1j * imag
    return out


def benchmark1d(args, data):
    if args.width % 2 != 1:
        raise SystemExit("Width must be odd")
    if data.shape[0] <= args.width:
        raise SystemExit("Channels cannot be less than the filter width")
    context = accel.create_some_context(True)
    command_queue = context.create_command_queue(profile=True)
    background = rfi_device.BackgroundMedianFilterDeviceTemplate(context, args.width)
    noise_est = rfi_device.NoiseEstMADTDeviceTemplate(context, max(10240, data.shape[0]))
    threshold = rfi_device.ThresholdSumDeviceTemplate(context, n_windows=args.windows)
    template = rfi_device.FlaggerDeviceTemplate(background, noise_est, threshold)
    flagger = template.instantiate(command_queue, data.shape[0], data.shape[1],
                                   threshold_args={"n_sigma": args.sigmas})
    flagger.ensure_all_bound()
    data_device = flagger.buffer("vis")
    flags_device = flagger.buffer("flags")
    data_device.set(command_queue, data)
    flagger()                 # warm-up
    command_queue.finish()

    start_time = time.time()
    start_event = command_queue.enqueue_marker()
    for _ in range(args.repeat):
        flagger()
    end_event = command_queue.enqueue_marker()
    command_queue.finish()
    end_time = time.time()
    flags = flags_device.get(command_queue)
    print("Host time (ms):  ", (end_time - start_time) * 1000.0 / args.repeat)
    device_ms = end_event.time_since(start_event) * 1000.0 / args.repeat
    print("Device time (ms):", device_ms)
    print(f"Throughput: {data.size / device_ms / 1e6:.2f} Gvis/s on {context.device.name}")
    return flags


def main():
    parser = argparse.ArgumentParser(description=__doc__)
    parser.add_argument("--antennas", "-a", type=int, default=7)
    parser.add_argument("--channels", "-c", type=int, default=1024)
    parser.add_argument("--baselines", "-b", type=int, help="(overrides --antennas)")
    parser.add_argument("--preset", "-p", choices=("small", "medium", "big", "kat7", "meerkat"),
                        help="(overrides other options)")
    parser.add_argument("--file", type=str, help="specify a real data file (.npy)")
    parser.add_argument("--width", "-w", type=int, default=13,
                        help="median filter kernel size (must be odd)")
    parser.add_argument("--sigmas", type=float, default=11.0, help="threshold for detecting RFI")
    parser.add_argument("--windows", type=int, default=4, help="sum-threshold window sizes 1..2^(n-1)")
    parser.add_argument("--repeat", type=int, default=10)
    args = parser.parse_args()

    if args.file is not None:
        if not args.file.endswith(".npy"):
            raise SystemExit("Don't know how to handle " + args.file)
        data = np.load(args.file)
    else:
        presets = {"small": (2, max(2 * args.width, 8)), "medium": (15, max(2 * args.width, 128)),
                   "kat7": (7, 8192), "big": (64, 10240), "meerkat": (64, 32768)}
        if args.preset is not None:
            args.baselines = None
            args.antennas, args.channels = presets[args.preset]
        if args.baselines is None:
            args.baselines = args.antennas * (args.antennas + 1) * 2   # 4 polarisations
        data = generate_data(args.channels, args.baselines)
    flags = benchmark1d(args, data)
    print(f"{100.0 * np.sum(flags != 0) / flags.size:.4f}% flagged")


if __name__ == "__main__":
    main()
