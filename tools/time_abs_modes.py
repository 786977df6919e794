"""Background kernel under both amplitude rules (developer tool, GPU only)."""
import os
import sys
from ctypes import c_void_p

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from katsdpsigproc_b200 import _capi  # noqa: E402

C, B = 32768, 2368
torch.manual_seed(1)
vis = torch.randn(C, B, 2, device="cuda")
dev_t = torch.empty(B, C, device="cuda")
S = c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: c_void_p(t.data_ptr())
for abs_mode, name in ((0, "numpy rule"), (1, "hypot rule")):
    def run():
        _capi.call("ksp_background_median_filter_t", S, p(vis), p(dev_t), None, C, B, B, C, 0, 13, 0, 0, abs_mode)
    run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        run()
    b.record()
    torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) / 5:.3f} ms for {B} baselines", flush=True)
