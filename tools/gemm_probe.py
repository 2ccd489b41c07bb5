"""Timing ablations of the tcgen05 GEMM at the encoder's shapes (B200, CUDA events, >L2 operands).

    python tools/gemm_probe.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sam_road_b200 import _lib  # noqa: E402

lib = _lib.load()
DEV = "cuda"
st = torch.cuda.current_stream().cuda_stream
M = 65536


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def f16_case(name, N, K, acts):
    A = torch.randn(M, K, device=DEV).half()
    W = (torch.randn(N, K, device=DEV) / K ** 0.5).half()
    bias = torch.randn(N, device=DEV)
    out = torch.empty(M, N, dtype=torch.float16, device=DEV)
    for label, act in acts:
        us = timeit(lambda: _lib.check(lib.samroad_op_gemm_f16(A.data_ptr(), K, W.data_ptr(), K, M, N, K,
                                                               bias.data_ptr(), act, out.data_ptr(), N, st), "g"))
        print(f"{name:10s} N={N:5d} K={K:5d} {label:16s} {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s")


def f32_case(name, N, K):
    A = torch.randn(M, K, device=DEV).half()
    W = (torch.randn(N, K, device=DEV) / K ** 0.5).half()
    bias = torch.randn(N, device=DEV)
    x = torch.randn(M, N, device=DEV)
    x0 = x.clone()
    ref = None
    for label, mode in (("tma reduce-add", 0), ("tma load+store", 4), ("regs", 2), ("1cta", 1)):
        lib.samroad_debug_disable_2cta_gemm(mode)
        y = x0.clone()
        _lib.check(lib.samroad_op_gemm_f32(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(),
                                           y.data_ptr(), None, 0, y.data_ptr(), N, st), "g")
        if ref is None:
            ref = y
        diff = (y - ref).abs().max().item()
        us = timeit(lambda: _lib.check(lib.samroad_op_gemm_f32(A.data_ptr(), K, W.data_ptr(), K, M, N, K,
                                                               bias.data_ptr(), x.data_ptr(), None, 0,
                                                               x.data_ptr(), N, st), "g"))
        gbs = (M * N * 8 + M * K * 2) / us / 1e3
        print(f"{name:10s} N={N:5d} K={K:5d} {label:16s} {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s  {gbs:6.0f} GB/s  maxdiff vs first {diff:.2e}")
    lib.samroad_debug_disable_2cta_gemm(0)


ACTS = (("none", 0), ("gelu packed", 1), ("none direct st", 0x40), ("none, no TMA st", 0x80), ("gelu, no TMA st", 0x81), ("skip epilogue", 100))
SKIP = ACTS[:1] + ACTS[4:5]
which = set(sys.argv[1:])


def want(name):
    return not which or name in which


if want("lin1"):
    f16_case("lin1", 3072, 768, ACTS)
if want("qkv"):
    f16_case("qkv", 2304, 768, SKIP)
if want("k3072"):
    f16_case("k3072", 3072, 3072, ACTS[:2] + ACTS[4:5])
if want("shapes"):      # main-loop ceilings of the fp32-epilogue shapes
    f16_case("lin2shape", 768, 3072, SKIP)
    f16_case("projshape", 768, 768, SKIP)
    f16_case("n1536", 1536, 3072, SKIP)
if want("proj"):
    f32_case("proj", 768, 768)
if want("lin2"):
    f32_case("lin2", 768, 3072)
