"""Diagnostic (not collected): a handful of Dense-layer launches for an `ncu --set full` capture."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "outdoor-nerf-depth_b200"))
from nerfpp_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda:0")
L.mip360_debug_set_pair_mode.argtypes = [ctypes.c_int]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for mode, M, N, K in ((1, 131072, 1024, 1024), (0, 131072, 1024, 1024), (0, 262144, 256, 256), (1, 262144, 256, 512)):
    L.mip360_debug_set_pair_mode(mode)
    a = (torch.randn(M, K, device=dev) * 0.5).half()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
    b = torch.zeros(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    for _ in range(2):
        _lib.check(L.mip360_dense_f16(a.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, 1, st), "dense")
    torch.cuda.synchronize()
