"""Does launch latency matter?  The whole forward (78 launches per step of 8 segments) timed as plain launches and as
one CUDA graph replay (torch.cuda.CUDAGraph around TFLocoformerMSS.forward), 10 steps each.

    python profiles/graph_forward.py [batch]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import VARIANT_D, SEG, make_mixture, make_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = dict(VARIANT_D)
model = make_state_dict(cfg).cuda().eval()
model.precision = "bf16"
mix = make_mixture(B, SEG).cuda()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(n):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n


with torch.no_grad():
    eager = model(mix)
    print(f"plain launches: {timeit(lambda: model(mix)):.2f} ms per step")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            model(mix)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        captured = model(mix)
    print(f"graph replay:   {timeit(graph.replay):.2f} ms per step")
    graph.replay()
    torch.cuda.synchronize()
    print("max |graph - plain| =", max(float((captured[k] - eager[k]).abs().max()) for k in eager))
