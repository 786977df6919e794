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
