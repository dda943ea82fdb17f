"""Diagnostic (not a test): times the tcgen05 field kernel for each weight-sharing cluster size and prints the
MMA-issuer cycle counters (total / waiting for E / waiting for A / waiting for weights). Run on the GPU box."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import nerfpp_oracle as O
from nerfpp_b200 import _lib, ops, FIELD_TC, FIELD_SIMT
from test_parity_gpu import make_models, relerr

dev = torch.device("cuda:0")
L = _lib.lib()
L.nerfpp_debug_set_tc_timers.argtypes = [ctypes.c_void_p]
L.nerfpp_debug_set_tc_cluster.argtypes = [ctypes.c_int]
L.nerfpp_debug_set_tc_flags.argtypes = [ctypes.c_int]
L.nerfpp_debug_set_tc_flags(int(os.environ.get('FLAGS', '0')))
nets = make_models([O.densify(O.make_params(), 5.0)])
net = nets[0].nerf_net
n, S = 4096, 192
rays = O.synthetic_rays(n, seed=0)
o, d = rays["ray_o"].to(dev), rays["ray_d"].to(dev)
far = ops.intersect_sphere(o, d)
g = torch.Generator().manual_seed(0)
fg_z = (torch.sort(torch.rand(n, S, generator=g), -1)[0]).to(dev) * far[:, None]
bg_z = torch.sort(torch.rand(n, S, generator=g), -1)[0].to(dev)
ref = {}
for is_bg, z, tensors, macs in ((0, fg_z, net.fg_net.tensors(), 593408), (1, bg_z, net.bg_net.tensors(), 604160)):
    packed = net._packed[is_bg].get(tensors, FIELD_TC)
    for cl in [int(x) for x in os.environ.get("CLUSTERS", "1,2,4").split(",")]:
        L.nerfpp_debug_set_tc_cluster(cl)
        dbg = torch.zeros(80 * 148, dtype=torch.int64, device=dev)
        L.nerfpp_debug_set_tc_timers(ctypes.c_void_p(dbg.data_ptr()) if os.environ.get('TIMERS', '1') == '1' else None)
        try:
            for _ in range(3):
                sig, rgb, dr = ops.field_forward(packed, is_bg, o, d, z, FIELD_TC)
            torch.cuda.synchronize()
        except Exception as e:
            print("cluster", cl, "FAILED", e); continue
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        reps = 10
        for _ in range(reps):
            sig, rgb, dr = ops.field_forward(packed, is_bg, o, d, z, FIELD_TC)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        extra = dbg.cpu().numpy()[8 * 148:10 * 148].reshape(-1, 2); ep = dbg.cpu().numpy()[10 * 148:14 * 148].reshape(-1, 4); pr = dbg.cpu().numpy()[14 * 148:30 * 148].reshape(-1, 16); tl = dbg.cpu().numpy()[32 * 148:72 * 148].reshape(-1, 40); t = dbg.cpu().numpy()[:8 * 148].reshape(-1, 8)
        t = t[t[:, 0] > 0] if (t[:, 0] > 0).any() else t[:1]
        print("   issue(incl A waits) %.0f  commits %.0f" % (extra[:, 0].mean(), extra[:, 1].mean()), " epi: ld-wait %.0f cvt %.0f st %.0f arrive %.0f" % tuple(ep.mean(0)))
        if pr[:, 0].any():
            q = pr[pr[:, 0] > 0]
            rel = (q - q[:, :1]).mean(0)
            print("   chain (cycles after ACC commit of layer 2): epi wake %.0f | epi arrive c0..3 %s | mma past AREADY c0..3 %s | layer 3 ACC commit issued %.0f | chunk 0: ld done %.0f cvt done %.0f st done %.0f" % (rel[1], rel[8:12].round(), rel[4:8].round(), rel[12], rel[2], rel[3], rel[13]))
        if tl[:, 0].any():
            q = tl[tl[:, 0] > 0].astype(np.float64)
            rel = (q - q[:, :1]).mean(0)
            print("   tile timeline (cycles after layer-0 ACC commit issue): mma commit issued L0..9 %s | next tile L0,L1 %s" % (rel[0:10].round(), rel[30:32].round()))
            print("       epilogue wake L0..9 %s" % rel[10:20].round())
            print("       epilogue done L0..9 %s" % rel[20:30].round())
        key = (is_bg,)
        if key not in ref:
            ref[key] = (sig.clone(), rgb.clone())
        es, er = relerr(sig.cpu().numpy(), ref[key][0].cpu().numpy()), relerr(rgb.cpu().numpy(), ref[key][1].cpu().numpy())
        print("bg=%d cluster=%d  %.3f ms  %.0f TFLOP/s  ctas=%d  mma-thread cycles: total %.0f  wait_E %.0f  wait_A %.0f  wait_W %.0f | emb total %.0f wait %.0f | epi warp0: total %.0f wait_acc %.0f  (vs cluster1: sigma %.1e rgb %.1e)"
              % (is_bg, cl, ms, 2 * macs * n * S / ms / 1e9, len(t), t[:, 0].mean(), t[:, 1].mean(), t[:, 2].mean(), t[:, 3].mean(), t[:, 4].mean(), t[:, 7].mean(), t[:, 5].mean(), t[:, 6].mean(), es, er))
    L.nerfpp_debug_set_tc_timers(None)
