"""Run bench.py's GPU arm against the SIMT-emulated library on a CPU-only box -- TEST INFRASTRUCTURE ONLY.

The build container has no GPU, and bench.py is the one script the driver runs unattended at the end of a round: this
harness lets its control flow (legs, sweep caps, the strong-scaling point with its head-room guard, the JSON line) run here
on tiny boxes.  torch's CUDA entry points are replaced by CPU stand-ins (tensors stay on the host, synchronize is a
no-op) and MSKC_LIB points at tests/emu/_build/libmskc.so.  Numbers printed under it mean nothing; only world size 1 (the
emulated library has no NCCL).

  python tests/emu/bench_under_emu.py --box 192 192 96 --strong-box 192 192 48 --steps 2 --warmup 1 --sweeps 4
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
os.environ["MSKC_EMU"] = "1"
os.environ.setdefault("MSKC_LIB", os.path.join(HERE, "_build", "libmskc.so"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.Tensor.pin_memory = lambda self, *a, **k: self
torch.Tensor.cuda = lambda self, *a, **k: self
_tensor, _empty, _full = torch.tensor, torch.empty, torch.full


def _host(fn):
    def f(*a, **k):
        k.pop("device", None)
        return fn(*a, **k)
    return f


torch.tensor, torch.empty, torch.full = _host(_tensor), _host(_empty), _host(_full)

sys.argv[0] = os.path.join(ROOT, "bench.py")
import bench  # noqa: E402

if __name__ == "__main__":
    bench.main()
