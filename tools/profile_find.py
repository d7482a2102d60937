"""One launch each of find (1 % and 50 % selected) and the single-row cumsum on 2^26 fp32, for `ncu` captures under gpurun."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from matx_b200 import ops as mx  # noqa: E402

ex = mx.CudaExecutor()
n = 1 << 26
x = torch.rand(n, device="cuda")
out = torch.empty(n, device="cuda")
nf = torch.zeros((), dtype=torch.int32, device="cuda")
iout = torch.empty(n, dtype=torch.int32, device="cuda")
for thr in (0.99, 0.5):
    for _ in range(2):
        mx.mtie(mx.make_tensor(out), mx.make_tensor(nf)).set(mx.find(mx.make_tensor(x), mx.GT(thr))).run(ex)
    ex.sync()
    print(ex.last_kernel(), nf.item())
    for _ in range(2):
        mx.mtie(mx.make_tensor(iout), mx.make_tensor(nf)).set(mx.find_idx(mx.make_tensor(x), mx.GT(thr))).run(ex)
    ex.sync()
    print(ex.last_kernel(), nf.item())
for _ in range(2):
    mx.make_tensor(out).set(mx.cumsum(mx.make_tensor(x))).run(ex)
ex.sync()
print(ex.last_kernel())
