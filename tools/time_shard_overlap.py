"""Small shards of a dump (BASELINE.json configs[4] on 4 / 8 GPUs), 16 dumps per block: one flagger
on one queue against two flaggers on two queues taking alternate dumps (developer tool, GPU only).

    python tools/time_shard_overlap.py [baselines ...]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from katsdpsigproc_b200 import _capi, accel, rfi  # noqa: E402
from katsdpsigproc_b200.rfi import device as rfi_device  # noqa: E402


def main():
    shards = [int(x) for x in sys.argv[1:]] or [1632, 3264, 6496]
    C, dumps = 32768, 16
    context = accel.create_some_context(interactive=False)
    queues = [context.create_command_queue() for _ in range(3)]
    background = rfi_device.BackgroundMedianFilterDeviceTemplate(context, 13)
    noise = rfi_device.NoiseEstMADTDeviceTemplate(context, 32768)
    threshold = rfi_device.ThresholdSumDeviceTemplate(context, n_windows=7)
    template = rfi_device.FlaggerDeviceTemplate(background, noise, threshold)
    targs = {"n_sigma": 11.0, "threshold_falloff": 1.2}
    for nb in shards:
        fls = []
        for q in queues:
            fl = template.instantiate(q, C, nb, threshold_args=targs)
            if fls:
                fl.bind(vis=fls[0].buffer("vis"))
            fl.ensure_all_bound()
            fls.append(fl)
        vis = fls[0].buffer("vis")
        stride = vis.padded_shape[1]
        torch.manual_seed(nb)
        gen = torch.zeros(C, stride, 2, device="cuda", dtype=torch.float32)
        gen[:, :nb].normal_()
        hit = torch.rand(C, nb, device="cuda") < (1.0 / 64.0)
        gen[:, :nb, 0] += hit * (torch.rand(C, nb, device="cuda") * 20.0 + 50.0)
        torch.cuda.synchronize()
        _capi.call("ksp_memcpy_async", vis.ptr, gen.data_ptr(), gen.numel() * 4, _capi.D2D, queues[0].stream)
        queues[0].finish()
        del gen, hit
        for n_q in (1, 2, 3):
            times = []
            for rep in range(5):
                for q in queues:
                    q.finish()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for d in range(dumps):
                    fls[d % n_q]()
                for q in queues[:n_q]:
                    q.finish()
                times.append(1e3 * (time.perf_counter() - t0) / dumps)
            times.sort()
            print(f"baselines {nb}: {n_q} queue(s) {times[len(times) // 2]:.4f} ms per dump (wall clock over {dumps} dumps)",
                  flush=True)
        flags = [np.asarray(fl.buffer("flags").get(queues[0])) for fl in fls]
        print("  flags equal across instances:", all(np.array_equal(flags[0], f) for f in flags[1:]), flush=True)
        del fls


if __name__ == "__main__":
    main()
