import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """-> (cfg dict, state_dict of torch tensors, dict of the other arrays as torch tensors)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = json.loads(bytes(z["config"]).decode())
    sd, arrays = {}, {}
    for k in z.files:
        if k == "config":
            continue
        t = torch.from_numpy(z[k])
        if k.startswith("sd/"):
            sd[k[3:]] = t
        else:
            arrays[k] = t
    return cfg, sd, arrays


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(autouse=True)
def _no_expired_waits(request):
    """After every GPU test: no bounded mbarrier wait of the tcgen05 kernels may have expired (an expired wait traps
    the kernel; the record lives in host memory, so this check is free and never touches the device)."""
    yield
    if request.node.get_closest_marker("gpu") is None:
        return
    from mss_tf_locoformer_b200.engine import debug_timeout
    rec = debug_timeout(reset=False)
    assert rec[0] == 0, f"a bounded mbarrier wait expired: block {rec[1]}, thread {rec[2]}, barrier 0x{rec[3]:x}, parity {rec[4]}"
