"""What caps the end-to-end (host -> device) rate when N ranks upload at once?  (developer tool)

    python tools/h2d_probe.py                               # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/h2d_probe.py              # N ranks, one per GPU, all copying together

Every rank copies the 2.18 GB staging buffer of one MeerKAT dump (32768 x 8320 complex64) to its GPU,
as StreamingFlagger does, with the pinned buffer allocated three ways (cudaHostAllocDefault,
WriteCombined, Portable), as ONE cudaMemcpyAsync or split into 8 copies on two streams, alone or
with the 0.27 GB flag download running the other way.  All ranks start together (barrier); the
table gives GB/s per rank (min / mean / max over ranks) and the aggregate.
"""
import ctypes
import json
import os
import time

import torch
import torch.distributed as dist

H2D, D2H = 1, 2
FLAGS = {"default": 0, "portable": 1, "write_combined": 4}


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    rt.cudaFreeHost.argtypes = [ctypes.c_void_p]
    n_up, n_down = 32768 * 8320 * 8, 32768 * 8320
    dev_up = torch.empty(n_up, dtype=torch.uint8, device="cuda")
    dev_down = torch.zeros(n_down, dtype=torch.uint8, device="cuda")
    s = [torch.cuda.Stream() for _ in range(3)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run(host_up, host_down, pieces, with_d2h, reps=4):
        best = 0.0
        for _ in range(reps):
            barrier()
            t0 = time.perf_counter()
            step = n_up // pieces
            for i in range(pieces):
                rt.cudaMemcpyAsync(dev_up.data_ptr() + i * step, host_up + i * step, step, H2D,
                                   ctypes.c_void_p(s[i % 2].cuda_stream))
            if with_d2h:
                rt.cudaMemcpyAsync(host_down, dev_down.data_ptr(), n_down, D2H, ctypes.c_void_p(s[2].cuda_stream))
            s[0].synchronize()
            s[1].synchronize()
            t_up = time.perf_counter() - t0
            torch.cuda.synchronize()
            best = max(best, n_up / t_up / 1e9)
        return best

    results = {}
    for name, flag in FLAGS.items():
        up, down = ctypes.c_void_p(), ctypes.c_void_p()
        assert rt.cudaHostAlloc(ctypes.byref(up), n_up, flag) == 0
        assert rt.cudaHostAlloc(ctypes.byref(down), n_down, 0) == 0
        ctypes.memset(up, 1, n_up)                     # first touch on this rank's CPU
        for pieces in (1, 8):
            for with_d2h in (False, True):
                key = f"{name}, {pieces} cop{'y' if pieces == 1 else 'ies'}{', + D2H' if with_d2h else ''}"
                results[key] = run(up.value, down.value, pieces, with_d2h)
        rt.cudaFreeHost(up)
        rt.cudaFreeHost(down)
    try:
        cpus = sorted(os.sched_getaffinity(0))
        aff = f"{cpus[0]}-{cpus[-1]} ({len(cpus)})"
    except Exception:
        aff = "?"
    mine = {"rank": rank, "gpu": torch.cuda.get_device_name(local), "cpu_affinity": aff, "GBps": results}
    if world > 1:
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
    else:
        everyone = [mine]
    if rank == 0:
        print(f"# H2D of one 2.18 GB dump per rank, {world} rank(s) copying together; GB/s per rank, best of 4")
        print(f"{'pinned allocation, copies':44s} {'min':>7s} {'mean':>7s} {'max':>7s} {'aggregate':>10s}")
        for key in results:
            v = [e["GBps"][key] for e in everyone]
            print(f"{key:44s} {min(v):7.1f} {sum(v) / len(v):7.1f} {max(v):7.1f} {sum(v):10.1f}")
        print("# cpu affinity per rank:", [e["cpu_affinity"] for e in everyone])
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/h2d_probe_n{world}.json", "w") as f:
            json.dump(everyone, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
