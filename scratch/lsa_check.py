import sys, time
sys.path.insert(0, '.')
import numpy as np, torch as th
from scipy.optimize import linear_sum_assignment
from marl_sap_b200 import _lib
lib = _lib.load()
def run(B, n, m, noise, seed):
    g = th.Generator(device='cuda').manual_seed(seed)
    q = th.randn(B, n, m, device='cuda', generator=g)
    z = th.randn(B, n, m, device='cuda', generator=g) if noise else None
    std = (q.abs().mean(dim=(1, 2)) * 0.3 * 2).contiguous() if noise else None
    out = th.empty(B, n, dtype=th.int64, device='cuda'); obj = th.empty(B, dtype=th.float64, device='cuda')
    def call():
        _lib.check(lib.sap_lsa_maximize(q.data_ptr(), _lib.ptr(z), _lib.ptr(std), B, n, m, out.data_ptr(), obj.data_ptr(), _lib.stream_ptr()), "lsa")
    call(); th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record(); th.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ben = q if not noise else q + z * std[:, None, None]
    bn = ben.cpu().numpy(); o = out.cpu().numpy()
    nb = min(B, 16); bad = 0; t0 = time.perf_counter()
    for b in range(nb):
        r, c = linear_sum_assignment(bn[b], maximize=True)
        want = bn[b][r, c].astype(np.float64).sum(); got = bn[b][np.arange(n), o[b]].astype(np.float64).sum()
        assert len(set(o[b].tolist())) == n and o[b].min() >= 0 and o[b].max() < m, "infeasible"
        if not np.array_equal(c, o[b]): bad += 1
        assert abs(want - got) <= 1e-9 * max(1, abs(want)), (want, got)
        assert abs(obj[b].item() - got) <= 1e-9 * max(1, abs(got))
    cpu_ms = (time.perf_counter() - t0) / nb * 1e3
    print(f"B={B} n={n} m={m} noise={noise}: kernel {ms:.3f} ms ({ms / B * 1e3:.2f} us/env), scipy {cpu_ms:.3f} ms/env, assignments differing {bad}/{nb}")
run(8, 4, 4, False, 0); run(64, 10, 10, True, 1); run(256, 50, 50, True, 2); run(4096, 100, 100, True, 3)
run(64, 37, 53, True, 4); run(64, 324, 450, True, 5); run(16, 200, 512, False, 6)
def hard(B, n, m, seed=9):
    g = th.Generator(device='cuda').manual_seed(seed)
    q = (th.randn(B, 1, m, device='cuda', generator=g) + 0.05 * th.randn(B, n, m, device='cuda', generator=g)).contiguous()
    z = th.randn(B, n, m, device='cuda', generator=g); std = (q.abs().mean(dim=(1, 2)) * 0.05 * 2).contiguous()
    out = th.empty(B, n, dtype=th.int64, device='cuda'); obj = th.empty(B, dtype=th.float64, device='cuda')
    def call():
        _lib.check(lib.sap_lsa_maximize(q.data_ptr(), z.data_ptr(), std.data_ptr(), B, n, m, out.data_ptr(), obj.data_ptr(), _lib.stream_ptr()), "lsa")
    call(); th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record(); th.cuda.synchronize()
    bn = (q + z * std[:, None, None]).cpu().numpy(); o = out.cpu().numpy(); same = 0
    for b in range(8):
        r, c = linear_sum_assignment(bn[b], maximize=True); same += int(np.array_equal(c, o[b]))
        assert abs(bn[b][r, c].astype(np.float64).sum() - obj[b].item()) < 1e-8
    print(f"hard B={B} n={n} m={m}: kernel {e0.elapsed_time(e1):.3f} ms, same assignment {same}/8")
hard(4096, 100, 100); hard(4096, 100, 120); hard(64, 324, 450)
