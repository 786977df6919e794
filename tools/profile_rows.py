"""Run each row-kernel variant a few times at cfg2 size (target for ncu; developer tool)."""
import os
import sys
from ctypes import byref, c_double, c_size_t, c_void_p

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from katsdpsigproc_b200 import _capi  # noqa: E402
import cabi_util as cu  # noqa: E402

C, B = 32768, int(sys.argv[1]) if len(sys.argv) > 1 else 2368
torch.manual_seed(1)
S = c_void_p(torch.cuda.current_stream().cuda_stream)
vis = torch.randn(C, B, 2, device="cuda")
spikes = torch.rand(C, B, device="cuda") < (1 / 64)
vis[..., 0] += spikes * (torch.rand(C, B, device="cuda") * 20 + 50)
dev_t = torch.empty(B, C, device="cuda")
noise = torch.empty(B, device="cuda")
flags_t = torch.empty(B, C, dtype=torch.uint8, device="cuda")
flags = torch.empty(C, B, dtype=torch.uint8, device="cuda")
p = lambda t: c_void_p(t.data_ptr())
sc7 = cu.scales(7, 1.2)
for _ in range(3):
    _capi.call("ksp_background_median_filter_t", S, p(vis), p(dev_t), None, C, B, B, C, 0, 13, 0, 0, 0)
    _capi.call("ksp_madnz_t", S, p(dev_t), p(noise), C, B, C)
    _capi.call("ksp_threshold_sum", S, p(dev_t), p(noise), p(flags_t), C, B, C, C, 7, c_double(11.0), sc7, 1)
    prm = cu.flagger_params(C, B, B, B, n_windows=7, chunk_baselines=B)
    nbytes = _capi.load().ksp_flagger_scratch_bytes(byref(prm))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    _capi.call("ksp_flagger", S, byref(prm), p(vis), None, p(noise), p(flags), p(scratch), c_size_t(nbytes))
    torch.cuda.synchronize()
print("ok", float(flags.float().mean()))
