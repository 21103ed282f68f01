"""SM cycles per synchronisation primitive (b200taco_measure_sync_costs)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pytorch-tacotron2-implementation_b200"))
import torch  # noqa: E402
from b200taco import _cabi  # noqa: E402

lib = _cabi.lib()
torch.zeros(1, device="cuda")
out = (C.c_int64 * 8)()
for _ in range(2):
    _cabi.check(lib.b200taco_measure_sync_costs(out, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "sync_costs")
names = ["fence.proxy.async", "fence.proxy.async.global", "fence.proxy.async.shared::cta", "__threadfence", "red.release.gpu",
         "ld.acquire.gpu", "atomicAdd(ret)", "st.global"]
for n, v in zip(names, out):
    print(f"{n:32s} {v:6d} cycles")
