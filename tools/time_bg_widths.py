"""Background filter timing for widths other than 13 (developer tool, GPU only)."""
import os
import sys
from ctypes import c_void_p

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from katsdpsigproc_b200 import _capi  # noqa: E402

C, B = 32768, int(sys.argv[1]) if len(sys.argv) > 1 else 2080
torch.manual_seed(1)
vis = torch.randn(C, B, 2, device="cuda")
dev_t = torch.empty(B, C, device="cuda")
S = c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: c_void_p(t.data_ptr())
for width in (3, 5, 9, 11, 13, 15, 21, 31):
    def run():
        _capi.call("ksp_background_median_filter_t", S, p(vis), p(dev_t), None, C, B, B, C, 0, width, 0, 0, 0)
    run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"width {width:2d}: {ms:8.3f} ms for {B} baselines = {C * B / ms / 1e6:8.1f} Gvis/s", flush=True)

# SumThreshold with more than 7 window sizes (general kernel)
import ctypes
from ctypes import c_double
noise = torch.full((B,), 1.0, device="cuda")
flags_t = torch.empty(B, C, dtype=torch.uint8, device="cuda")
dev_t.normal_()
for n_windows in (7, 8, 10, 11):
    scales = (c_double * n_windows)(*[1.2 ** -i for i in range(n_windows)])
    def run():
        _capi.call("ksp_threshold_sum", S, p(dev_t), p(noise), p(flags_t), C, B, C, C, n_windows,
                   c_double(11.0), scales, 1)
    run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"threshold_sum n_windows {n_windows:2d}: {ms:8.3f} ms for {B} baselines = {C * B / ms / 1e6:8.1f} Gvis/s",
          flush=True)
