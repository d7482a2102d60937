"""One launch each of the one-pass variance kernels (full tensor, 256 KB rows, column variance through the TMA tiles),
for `ncu` captures under gpurun.  Not a benchmark."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from matx_b200 import ops as mx  # noqa: E402

ex = mx.CudaExecutor()
x = torch.rand(1 << 30, device="cuda") + 0.5
o = torch.empty((), device="cuda")
mx.make_tensor(o).set(mx.var(mx.make_tensor(x), None, 1)).run(ex)
ex.sync()
print(ex.last_kernel())
del x
y = torch.rand(4096, 65536, device="cuda") + 0.5
for dims, n in (([1], 4096), ([0], 65536)):
    o = torch.empty(n, device="cuda")
    mx.make_tensor(o).set(mx.var(mx.make_tensor(y), dims, 1)).run(ex)
    ex.sync()
    print(ex.last_kernel())
