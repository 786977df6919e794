"""Per-kernel device timings at the BASELINE.json cfg2 shape (developer tool, GPU only).

Usage: python tools/time_kernels.py [channels baselines] [--reps N]
Inputs are generated on the device with torch (plumbing only); every timed call
goes through the C ABI on torch's current stream.
"""
import argparse
import ctypes
import json
import os
import sys
from ctypes import byref, c_double, c_size_t, c_void_p

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from katsdpsigproc_b200 import _capi  # noqa: E402
import cabi_util as cu  # noqa: E402


def timeit(fn, reps, flush):
    fn()
    torch.cuda.synchronize()
    if reps <= 0:               # one launch only (under ncu)
        return float("nan")
    times = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    times.sort()
    return times[len(times) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("channels", type=int, nargs="?", default=32768)
    ap.add_argument("baselines", type=int, nargs="?", default=8320)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--chunks", type=str, default="0")
    ap.add_argument("--only-fused", action="store_true")
    ap.add_argument("--sweep", action="store_true",
                    help="also the helper configurations of BASELINE.json configs[2], [3]: Percentile5 "
                         "over 8320 rows x 256..65536 columns, complex64 transpose 3072 x 8320")
    ap.add_argument("--dataflow", action="store_true", help="also the dataflow form of the flagger")
    ap.add_argument("--tpad", type=int, default=0,
                    help="extra elements in the row stride of the baseline-major arrays")
    args = ap.parse_args()
    C, B = args.channels, args.baselines
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    S = c_void_p(torch.cuda.current_stream().cuda_stream)
    vis = torch.randn(C, B, 2, device=dev, dtype=torch.float32)
    spikes = torch.rand(C, B, device=dev) < (1 / 64)
    vis[..., 0] += spikes * (torch.rand(C, B, device=dev) * 20 + 50)
    del spikes
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    dev_cm = torch.empty(C, B, dtype=torch.float32, device=dev)
    CT = C + args.tpad                      # row stride of dev_t / flags_t
    dev_t = torch.empty(B, CT, dtype=torch.float32, device=dev)
    noise = torch.empty(B, dtype=torch.float32, device=dev)
    flags_t = torch.empty(B, CT, dtype=torch.uint8, device=dev)
    flags = torch.empty(C, B, dtype=torch.uint8, device=dev)
    N = C * B
    res = {}
    p = lambda t: c_void_p(t.data_ptr())
    sc7 = cu.scales(7, 1.2)

    def rec(name, ms, bytes_per_vis):
        res[name] = {"ms": round(ms, 4), "GB/s": round(bytes_per_vis * N / ms / 1e6, 1)}
        print(name, res[name], flush=True)

    if not args.only_fused:
        rec("background", timeit(lambda: _capi.call(
            "ksp_background_median_filter", S, p(vis), p(dev_cm), None, C, B, B, B, 0, 13, 0, 0, 0),
            args.reps, flush), 12)
        rec("background_t", timeit(lambda: _capi.call(
            "ksp_background_median_filter_t", S, p(vis), p(dev_t), None, C, B, B, CT, 0, 13, 0, 0, 0),
            args.reps, flush), 12)
        amp = vis[..., 0].abs().contiguous()
        rec("background_t_amplitude_input", timeit(lambda: _capi.call(
            "ksp_background_median_filter_t", S, p(amp), p(dev_t), None, C, B, B, CT, 0, 13, 1, 0, 0),
            args.reps, flush), 8)
        del amp
        rec("transpose_f32", timeit(lambda: _capi.call(
            "ksp_transpose", S, p(dev_t), p(dev_cm), C, B, CT, B, 4), args.reps, flush), 8)
        rec("madnz_t", timeit(lambda: _capi.call(
            "ksp_madnz_t", S, p(dev_t), p(noise), C, B, CT), args.reps, flush), 4)
        fb = ctypes.c_ulonglong(0)
        _capi.call("ksp_selection_fallback_count", S, byref(fb), 1)
        _capi.call("ksp_madnz_t", S, p(dev_t), p(noise), C, B, CT)
        _capi.call("ksp_selection_fallback_count", S, byref(fb), 1)
        print("madnz_t fallback rows", fb.value, "of", B, flush=True)
        rec("madnz", timeit(lambda: _capi.call(
            "ksp_madnz", S, p(dev_cm), p(noise), C, B, B), args.reps, flush), 4)
        # what NoiseEstMADDevice runs for long rows: transposition into its scratch, then the row kernel
        def madnz_via_transpose():
            _capi.call("ksp_transpose", S, p(dev_t), p(dev_cm), C, B, CT, B, 4)
            _capi.call("ksp_madnz_t", S, p(dev_t), p(noise), C, B, CT)
        rec("madnz_via_transpose", timeit(madnz_via_transpose, args.reps, flush), 4)
        rec("threshold_sum7", timeit(lambda: _capi.call(
            "ksp_threshold_sum", S, p(dev_t), p(noise), p(flags_t), C, B, CT, CT, 7, c_double(11.0), sc7, 1),
            args.reps, flush), 5)
        rec("threshold_simple_t", timeit(lambda: _capi.call(
            "ksp_threshold_simple", S, p(dev_t), p(noise), p(flags_t), B, C, CT, CT, c_double(11.0), 1, 1),
            args.reps, flush), 5)
        rec("transpose_u8", timeit(lambda: _capi.call(
            "ksp_transpose", S, p(flags), p(flags_t), B, C, B, CT, 1), args.reps, flush), 2)
    print("flagged fraction", float(flags.float().mean()))
    for chunk in [int(x) for x in args.chunks.split(",")] + ([-1] if args.dataflow else []):
        prm = cu.flagger_params(C, B, B, B, n_windows=7, chunk_baselines=chunk)
        nbytes = _capi.load().ksp_flagger_scratch_bytes(byref(prm))
        used = _capi.load().ksp_flagger_chunk_baselines(byref(prm))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        dataflow = bool(_capi.load().ksp_flagger_is_dataflow(byref(prm)))
        flags.fill_(0x55)
        name = "flagger_dataflow" if dataflow else f"flagger_fused_chunk{used}"
        rec(name, timeit(lambda: _capi.call(
            "ksp_flagger", S, byref(prm), p(vis), None, p(noise), p(flags), p(scratch),
            c_size_t(nbytes)), args.reps, flush), 9)
        st = (ctypes.c_ulonglong * len(_capi.DF_STAT_NAMES))()
        _capi.call("ksp_flagger_stats", S, byref(prm), p(scratch), st, len(st))
        res[name]["scratch_MB"] = round(nbytes / 1e6, 1)
        res[name]["flags_sum"] = int(flags.sum(dtype=torch.int64))
        res[name]["noise_sum"] = float(noise.double().sum())
        if dataflow:
            res[name]["stats"] = dict(zip(_capi.DF_STAT_NAMES, (int(v) for v in st)))
        print(name, res[name], flush=True)
        del scratch
    print("flagged fraction (fused)", float(flags.float().mean()))
    if not args.only_fused:
        pct = torch.empty(5, B, dtype=torch.float32, device=dev)
        rec("percentile5_f32", timeit(lambda: _capi.call(
            "ksp_percentile5", S, p(dev_t), p(pct), B, CT, B, 0, C, 1, 0), args.reps, flush), 4)
        # complex input: rows of the channel-major dump, as the reference's percentile script does
        pct_c = torch.empty(5, C, dtype=torch.float32, device=dev)
        rec("percentile5_c64", timeit(lambda: _capi.call(
            "ksp_percentile5", S, p(vis), p(pct_c), C, B, C, 0, B, 0, 0), args.reps, flush), 8)
        mask = (torch.rand(C, device=dev) < 0.9).float()
        dest = torch.empty(B, 2, dtype=torch.float32, device=dev)
        rec("maskedsum_c64", timeit(lambda: _capi.call(
            "ksp_maskedsum", S, p(vis), p(mask), p(dest), C, B, B, 0, 0), args.reps, flush), 8)
    if args.sweep:
        del dev_cm, dev_t, flags_t
        torch.cuda.empty_cache()
        rows = B
        for cols in (256, 1024, 4096, 16384, 65536):
            src = torch.randn(rows, cols, device=dev, dtype=torch.float32).abs_()
            pct = torch.empty(5, rows, dtype=torch.float32, device=dev)
            ms = timeit(lambda: _capi.call("ksp_percentile5", S, p(src), p(pct), rows, cols, rows, 0, cols, 1, 0),
                        args.reps, flush)
            res[f"percentile5_f32_{rows}x{cols}"] = {"ms": round(ms, 4),
                                                     "GB/s": round(4 * rows * cols / ms / 1e6, 1)}
            print(f"percentile5_f32_{rows}x{cols}", res[f"percentile5_f32_{rows}x{cols}"], flush=True)
            ms = timeit(lambda: _capi.call("ksp_madnz_t", S, p(src), p(pct), cols, rows, cols), args.reps, flush)
            res[f"madnz_t_{rows}x{cols}"] = {"ms": round(ms, 4), "GB/s": round(4 * rows * cols / ms / 1e6, 1)}
            print(f"madnz_t_{rows}x{cols}", res[f"madnz_t_{rows}x{cols}"], flush=True)
            del src, pct
        a64 = torch.randn(3072, B, 2, device=dev, dtype=torch.float32)
        b64 = torch.empty(B, 3072, 2, device=dev, dtype=torch.float32)
        ms = timeit(lambda: _capi.call("ksp_transpose", S, p(b64), p(a64), 3072, B, 3072, B, 8), args.reps, flush)
        res["transpose_c64_3072x8320"] = {"ms": round(ms, 4), "GB/s": round(16 * 3072 * B / ms / 1e6, 1)}
        print("transpose_c64_3072x8320", res["transpose_c64_3072x8320"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.environ.get("TK_OUT", "gpurun_out/time_kernels.json"), "w") as f:
        json.dump({"channels": C, "baselines": B, "results": res}, f, indent=1)


if __name__ == "__main__":
    main()
