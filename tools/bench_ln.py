"""LayerNorm fwd/bwd microbenchmark at the BERT-large shape: algorithmic bytes / time vs the measured HBM peak.
DLE_LN_ONE_WARP=1 selects the one-warp-per-row kernels for an A/B comparison."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearningexamples_b200 import kernels as k


def timeit(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


T, H = int(os.environ.get("T", 32768)), 1024
bf = torch.bfloat16
zs = [torch.randn(T, H, device="cuda").to(bf) for _ in range(4)]     # rotate inputs (> L2)
dys = [torch.randn(T, H, device="cuda").to(bf) for _ in range(4)]
g, b = torch.ones(H, device="cuda", dtype=bf), torch.zeros(H, device="cuda", dtype=bf)
y, z, mean, rstd = k.add_ln_fwd(zs[0], g, b)
it = [0]
def fwd():
    it[0] += 1; k.add_ln_fwd(zs[it[0] % 4], g, b)
def bwd(p):
    it[0] += 1; k.add_ln_bwd(dys[it[0] % 4], zs[it[0] % 4], mean, rstd, g, dropout_p=p, seed=1, dropout_stream=1)
res = {}
for name, fn, nbytes in (("fwd (read z, write y)", fwd, 4 * T * H), ("bwd p=0 (read dy,z; write dz)", lambda: bwd(0.0), 6 * T * H),
                         ("bwd p=0.1 (read dy,z; write dz,dx)", lambda: bwd(0.1), 8 * T * H)):
    ms = timeit(fn)
    res[name] = dict(us=round(ms * 1e3, 1), GBps=round(nbytes / ms / 1e6, 1))
    print(os.environ.get("DLE_LN_ONE_WARP", "0"), name, res[name], flush=True)
