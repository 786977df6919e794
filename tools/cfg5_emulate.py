"""BASELINE.json configs[4] on ONE GPU, shard by shard (developer tool, GPU only): the cfg5 block
of bench.py as rank 0 of 2, 4 and 8 ranks - the ranks are independent (no collective on the data
path), so the device-resident time of a shard is what each rank of the real run measures; the real
run is `torchrun ... bench.py --gpus N` (SCALE_rNN.json).

    python tools/cfg5_emulate.py [world ...]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from katsdpsigproc_b200 import cuda  # noqa: E402
from katsdpsigproc_b200.rfi import device as rfi_device  # noqa: E402


def main():
    worlds = [int(x) for x in sys.argv[1:]] or [2, 4, 8]
    args = argparse.Namespace(channels=bench.CHANNELS)
    context = cuda.Device(0).make_context()
    queue = context.create_command_queue()
    template = rfi_device.FlaggerDeviceTemplate(
        rfi_device.BackgroundMedianFilterDeviceTemplate(context, bench.WIDTH),
        rfi_device.NoiseEstMADTDeviceTemplate(context, max(args.channels, 10240)),
        rfi_device.ThresholdSumDeviceTemplate(context, n_windows=bench.N_WINDOWS))
    for world in worlds:
        out = bench.cfg5_block(template, queue, context, 0, world, torch.cuda.synchronize, lambda x: x, args)
        out.pop("e2e", None)
        out["emulated"] = f"rank 0 of {world} on one GPU"
        print(json.dumps(out), flush=True)
        print(f"SUMMARY world {world} shard {out['baselines_this_rank']} KSP_CHUNK={os.environ.get('KSP_CHUNK', '-')} "
              f"KSP_BG_WHOLE={os.environ.get('KSP_BG_WHOLE', '-')}: {out['ms_per_dump']:.4f} ms per dump, "
              f"whole dump {out['ms_per_dump_whole_dump_on_one_gpu']:.4f}, efficiency "
              f"{out['strong_scaling_efficiency']:.3f}, {out['launches_per_dump']} launches", flush=True)


if __name__ == "__main__":
    main()
