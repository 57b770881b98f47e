import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import mss_tf_locoformer_b200 as pkg
from test_gpu_parity import VARIANT_D, _random_model
model = _random_model(pkg, VARIANT_D).cuda().eval()
eng = model._ready()
x = torch.randn(1, 5, 300, 128, device='cuda')
try:
    eng.ffn_(0, 0, 0, x, 1)
    torch.cuda.synchronize()
    print("ok")
    from mss_tf_locoformer_b200.engine import debug_timeout
    print("timeout info", debug_timeout())
except Exception as e:
    print("ERR", str(e)[:200])
