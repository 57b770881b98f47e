"""Pipeline event trace of the tcgen05 FFN kernel (block 0): prints per-chunk timing in SM clocks."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import VARIANT_D, SEG, make_state_dict
from mss_tf_locoformer_b200 import _lib

cfg = dict(VARIANT_D)
model = make_state_dict(cfg).cuda()
eng = model._ready()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
BASE = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # first chunk recorded (steady state: e.g. 300)
Tf, F = 1 + SEG // cfg["hop_length"], cfg["n_fft"] // 2 + 1
x = torch.randn(B, Tf, F, cfg["emb_dim"], device="cuda")
for _ in range(2):
    eng.ffn_(0, 0, 0, x, 1)
buf = torch.zeros(16 * 64, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.tfl_debug_set_option(2, BASE)
lib.tfl_debug_set_trace(buf.data_ptr())
eng.ffn_(0, 0, 0, x, 1)
torch.cuda.synchronize()
lib.tfl_debug_set_trace(None)
t = buf.cpu().view(16, 64)
names = {0: "m1_start", 1: "m1_ready", 2: "m1_issued", 3: "epi_d1full", 4: "epi_d1empty", 10: "epi_math_done", 5: "epi_gempty",
         6: "epi_gfull", 7: "m2_start", 8: "m2_gfull", 9: "m2_issued"}
base = int(t[0, 6])
print("chunk " + " ".join(f"{names[e]:>13s}" for e in (0, 1, 2, 3, 4, 10, 5, 6, 7, 8, 9)))
print("(last two columns: clocks the MMA warp spent waiting for weight stages in M1 / M2 of the chunk)")
def rel(v):
    return int(v) - base if int(v) else -1


print("(last three columns: clocks the MMA warp waited for weight stages in M1 / for the peer's half / for the A tile)")
for c in range(6, 40):
    print(f"{c + BASE:5d} " + " ".join(f"{rel(t[e, c]):13d}" for e in (0, 1, 2, 3, 4, 10, 5, 6, 7, 8, 9)) +
          f" {int(t[11, c]):8d} {int(t[12, c]):8d} {int(t[15, c]):8d}")

print("output warp 16: output pass start (D2 full seen), end")
for it in range(1, 8):
    print(it + BASE, int(t[13, it]) - base, int(t[14, it]) - base)
