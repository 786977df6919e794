"""Experiment: replay the whole flagger call as a CUDA graph for small (L2-sized) chunks."""
import os
import sys
from ctypes import byref, c_size_t, c_void_p

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from katsdpsigproc_b200 import _capi  # noqa: E402
import cabi_util as cu  # noqa: E402

C, B = 32768, 8320
torch.manual_seed(1)
vis = torch.randn(C, B, 2, device="cuda")
spikes = torch.rand(C, B, device="cuda") < (1 / 64)
vis[..., 0] += spikes * (torch.rand(C, B, device="cuda") * 20 + 50)
del spikes
noise = torch.empty(B, device="cuda")
flags = torch.empty(C, B, dtype=torch.uint8, device="cuda")
p = lambda t: c_void_p(t.data_ptr())

for lanes in (4, 8):
    os.environ["KSP_LANES"] = str(lanes)
    for chunk in (96, 192, 296, 592, 2368):
        prm = cu.flagger_params(C, B, B, B, n_windows=7, chunk_baselines=chunk)
        nbytes = _capi.load().ksp_flagger_scratch_bytes(byref(prm))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            S = c_void_p(side.cuda_stream)
            for _ in range(2):
                _capi.call("ksp_flagger", S, byref(prm), p(vis), None, p(noise), p(flags), p(scratch), c_size_t(nbytes))
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            S = c_void_p(torch.cuda.current_stream().cuda_stream)
            _capi.call("ksp_flagger", S, byref(prm), p(vis), None, p(noise), p(flags), p(scratch), c_size_t(nbytes))
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        print(f"lanes {lanes} chunk {chunk:5d}: {a.elapsed_time(b) / 10:.4f} ms (graph)", flush=True)
        del g, scratch
