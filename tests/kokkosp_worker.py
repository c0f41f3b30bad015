"""Runs a few Kokkos user calls on Kokkos::B200 and Kokkos::Cuda (benchlib/libkokkos_arms.so) with a KokkosP tool loaded through
KOKKOS_TOOLS_LIBS (set by tests/test_gpu_adapter.py before this process starts).  Prints "hooks done"; the tool writes its log."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchlib import arms as A  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
arms = A.Arms(0, stream.cuda_stream)
n = 1 << 16
x = torch.ones(n, dtype=torch.float64, device=dev)
xi = torch.ones(n, dtype=torch.int64, device=dev)
y = torch.empty_like(xi)
a = torch.empty_like(x)
r = torch.zeros(1, dtype=torch.float64, device=dev)
t = torch.zeros(1, dtype=torch.int64, device=dev)
for arm in (A.B200, A.CUDA):
    arms.stream_copy(arm, x.data_ptr(), a.data_ptr(), n)
    arms.reduce_sum(arm, x.data_ptr(), n, r.data_ptr())
    arms.scan_excl(arm, xi.data_ptr(), y.data_ptr(), n, t.data_ptr())
torch.cuda.synchronize()
assert float(r.item()) == n and int(t.item()) == n
arms.finalize()
print("hooks done")
