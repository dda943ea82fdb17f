cd /root/repo
timeout 1500 python -m pytest tests/test_convergence_gpu.py -q -m gpu -s 2>&1 | grep -v "Warning\|warn" > gpurun_out/exp6_conv.log; grep -n "PSNR oracle\|^split\|^rgb \|^depth\|^fg_depth\|passed\|failed\|Error" gpurun_out/exp6_conv.log | head -40
echo "== parity tests (existing)"
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -x 2>&1 | tail -3
echo "== split precision rate"
timeout 300 python - <<'PY'
import sys, torch
sys.path.insert(0,'tests'); import conftest
import nerfpp_oracle as O
from test_parity_gpu import make_models
from nerfpp_b200 import ops, FIELD_TC, FIELD_TC_SPLIT, FIELD_SIMT
net = make_models([O.densify(O.make_params(),5.0)])[0].nerf_net
rays = {k:(v.cuda() if torch.is_tensor(v) else v) for k,v in O.synthetic_rays(4096,seed=0).items()}
far = ops.intersect_sphere(rays['ray_o'],rays['ray_d'])
z = torch.sort(torch.rand(4096,192,device='cuda'),-1)[0]*far[:,None]
for name,impl in (('tc',FIELD_TC),('split',FIELD_TC_SPLIT)):
    pk = net._packed[0].get(net.fg_net.tensors(), impl)
    for _ in range(3): ops.field_forward(pk,0,rays['ray_o'],rays['ray_d'],z,impl)
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): s,r,_=ops.field_forward(pk,0,rays['ray_o'],rays['ray_d'],z,impl)
    b.record(); torch.cuda.synchronize()
    print(name, 'fg 4096x192: %.3f ms'%(a.elapsed_time(b)/10))
PY
