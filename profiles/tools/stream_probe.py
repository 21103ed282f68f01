"""Per-SM L2 streaming rate for LDG vs TMA bulk copies (b200taco_measure_stream)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pytorch-tacotron2-implementation_b200"))
import torch  # noqa: E402
from b200taco import _cabi  # noqa: E402

lib = _cabi.lib()
per_cta = 384 * 1024
buf = torch.zeros(148 * per_cta, dtype=torch.uint8, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for ctas in (1, 128):
    for mode, chunk, depth in ((4, 16384, 2), (4, 16384, -2), (8, 8192, -2), (4, 8192, -2), (4, 12288, -2)):
        if mode and per_cta % chunk:
            pc = per_cta // chunk * chunk
        else:
            pc = per_cta
        out = C.c_float(0)
        _cabi.check(lib.b200taco_measure_stream(buf.data_ptr(), pc, ctas, 200, mode, chunk, depth, C.byref(out), st), "stream")
        print(f"ctas {ctas:3d} mode {'LDG ' if mode == 0 else 'BULKx%d' % mode} chunk {chunk:6d} depth {depth:2d}: {out.value:7.1f} GB/s per SM, {out.value * ctas / 1e3:6.2f} TB/s total")
