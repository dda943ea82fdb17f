"""Diagnostic (not collected by pytest): the tcgen05 Dense layer of config 3 against torch, with timings."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "outdoor-nerf-depth_b200"))
from nerfpp_b200 import _lib
from nerfpp_b200._lib import check

L = _lib.lib()
dev = torch.device("cuda:0")
torch.manual_seed(0)


def run(M, N, K, relu, reps=0):
    a = (torch.randn(M, K, device=dev) * 0.5).half()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
    b = torch.randn(N, device=dev)
    out = torch.zeros(M, N, device=dev, dtype=torch.float16)
    st = torch.cuda.current_stream().cuda_stream
    check(L.mip360_dense_f16(a.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, relu, st), "dense")
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + b
    if relu:
        ref = ref.relu()
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    msg = "M %7d N %5d K %5d relu %d  max err %.3e (ref max %.2f)" % (M, N, K, relu, err, scale)
    if reps:
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        for _ in range(3):
            L.mip360_dense_f16(a.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, relu, st)
        e0.record()
        for _ in range(reps):
            L.mip360_dense_f16(a.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, relu, st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        msg += "  %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * M * N * K / ms / 1e9)
    print(msg, flush=True)
    return err, scale


import ctypes
for mode in (0, 1):
  L.mip360_debug_set_pair_mode.argtypes = [ctypes.c_int]
  L.mip360_debug_set_pair_mode(mode)
  print("---- pair mode", mode, flush=True)
  for shape in [(128, 128, 64, 0), (128, 256, 64, 1), (1000, 256, 512, 1), (333, 128, 320, 1), (4096, 1024, 1024, 1), (5000, 1024, 1536, 0)]:
    run(*shape)
  run(131072, 1024, 1024, 1, reps=10)
  run(131072, 1024, 1536, 1, reps=10)
  run(131072, 1024, 512, 1, reps=10)
  run(262144, 256, 256, 1, reps=10)
  run(262144, 256, 512, 1, reps=10)
  run(131072, 256, 1024, 0, reps=10)
  run(131072, 128, 320, 1, reps=10)
