"""Config 5 (bf16 1024^3, sum over the permuted dim) and two fp32 column-reduction shapes through the TMA-staged
reduce_outer family: strip width, stage size, ring depth, CTAs per SM, block size — with the LDG walker beside it.
Development tool, run under gpurun; every line also checks the result against fp32 device truth."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from matx_b200 import bench_configs as bc, ops as mx

ex = mx.CudaExecutor()
PEAK = 6456.8


def sweep(name, build, nbytes, check, envs):
    for env in envs:
        for k, v in env.items():
            os.environ[k] = str(v)
        try:
            ms, best = bc._time(ex, build, iters=8, warm=3)
            print(json.dumps({"case": name, "env": env, "ms": round(ms, 4), "best": round(best, 4), "GBps": round(nbytes / ms / 1e6, 1),
                              "frac": round(nbytes / ms / 1e6 / PEAK, 3), "ok": check(), "kernel": ex.last_kernel()}), flush=True)
        except Exception as exc:  # noqa: BLE001
            print(json.dumps({"case": name, "env": env, "error": str(exc)[:200]}), flush=True)
        for k in env:
            os.environ.pop(k)


ENVS = [{"MXB_OUTER_TMA": 0}, {}]
for mode in (1, 0):
    for tx in (32, 64, 128):
        for kb in (16, 32):
            ENVS.append({"MXB_OUTER_TMA_MODE": mode, "MXB_TUNE_TX": tx, "MXB_TUNE_OT_STAGE_KB": kb})
for tx in (64, 128):
    ENVS += [{"MXB_TUNE_TX": tx, "MXB_TUNE_OT_CTAS": 1, "MXB_TUNE_STAGES": 6}, {"MXB_TUNE_TX": tx, "MXB_TUNE_OT_CTAS": 1, "MXB_TUNE_STAGES": 6, "MXB_TUNE_BLOCK": 512},
             {"MXB_TUNE_TX": tx, "MXB_TUNE_OT_CTAS": 1, "MXB_TUNE_STAGES": 4, "MXB_TUNE_OT_STAGE_KB": 48, "MXB_TUNE_BLOCK": 512},
             {"MXB_TUNE_TX": tx, "MXB_TUNE_OT_CTAS": 3, "MXB_TUNE_OT_STAGE_KB": 16}, {"MXB_TUNE_TX": tx, "MXB_TUNE_BLOCK": 512},
             {"MXB_TUNE_TX": tx, "MXB_TUNE_STAGES": 2}, {"MXB_TUNE_TX": tx, "MXB_TUNE_OT_L2PROMO": 0}, {"MXB_TUNE_TX": tx, "MXB_TUNE_OT_L2PROMO": 3},
             {"MXB_TUNE_TX": tx, "MXB_TUNE_OT_STAGE_KB": 64, "MXB_TUNE_OT_CTAS": 1, "MXB_TUNE_STAGES": 3, "MXB_TUNE_BLOCK": 512}]
SHORT = ENVS[:2] + [{"MXB_OUTER_TMA_MODE": m, "MXB_TUNE_TX": tx} for m in (1, 0) for tx in (32, 64, 128)]

d = 1024
t = (torch.rand(d, d, d, device="cuda") * 0.25).to(torch.bfloat16)
out = torch.empty(d, d, dtype=torch.bfloat16, device="cuda")
tt, to = mx.make_tensor(t), mx.make_tensor(out)
want = t[:16].float().sum(1).t()


def check5():
    return bool((((out[:, :16].float() - want).abs() / want).max() <= 2 ** -8).item())


sweep("c5 bf16 1024^3 permuted sum", lambda: to.set(mx.sum(mx.permute(tt, [2, 0, 1]), [2])).run(ex), d * d * d * 2 + d * d * 2, check5, ENVS)
del t, out, tt, to
torch.cuda.empty_cache()

# fp32, same access pattern: [512, 1024, 1024] reduced over the middle dim (max: exact check)
x = torch.rand(512, 1024, 1024, device="cuda")
o = torch.empty(1024, 512, device="cuda")
tx_, to_ = mx.make_tensor(x), mx.make_tensor(o)
wantx = x[:8].amax(1).t()
sweep("fp32 512x1024x1024 permuted max", lambda: to_.set(mx.max(mx.permute(tx_, [2, 0, 1]), [2])).run(ex), x.numel() * 4 + o.numel() * 4,
      lambda: bool(torch.equal(o[:, :8], wantx)), SHORT)
del x, o
torch.cuda.empty_cache()

# column sums of a wide matrix: rows of 256 KB, one bulk copy per row of a strip
y = torch.rand(4096, 65536, device="cuda")
oc = torch.empty(65536, device="cuda")
ty_, toc = mx.make_tensor(y), mx.make_tensor(oc)
wanty = y.double().sum(0)
sweep("fp32 4096x65536 column sums", lambda: toc.set(mx.sum(ty_, [0])).run(ex), y.numel() * 4 + oc.numel() * 4,
      lambda: bool((((oc.double() - wanty).abs() / wanty).max() <= 1e-5).item()), SHORT)
