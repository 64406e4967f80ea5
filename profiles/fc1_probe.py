"""Timing probe for the agent's first layer at the bench shape: fp32 sgemm vs split-precision fp16 tensor-core GEMMs."""
import torch as th

rows, K, H = 409600, 490, 64
dev = "cuda"
x16 = th.rand(rows, K, device=dev).half()  # observations are fp16 values
x32 = x16.float()
W = (th.rand(H, K, device=dev) - 0.5) * 0.09
wt = W.t().contiguous()


def split(W, terms):
    parts, r = [], W.clone()
    for _ in range(terms):
        p = r.half()
        parts.append(p)
        r = (r - p.float()) * 2048.0
    return parts


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    th.cuda.synchronize()
    return a.elapsed_time(b) / reps


print("fp32 mm", timeit(lambda: th.mm(x32, wt)))
ref = th.mm(x32.double(), wt.double())
for terms in (2, 3):
    parts = split(W, terms)
    Wcat = th.cat([p.t() for p in parts], dim=1).contiguous()
    y = th.mm(x16, Wcat, out_dtype=th.float32)
    out = sum(y[:, k * H:(k + 1) * H] * (2.0 ** (-11 * k)) for k in range(terms))
    err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    e32 = ((th.mm(x32, wt).double() - ref).abs().max() / ref.abs().max()).item()
    print(f"terms {terms}: fp16 mm lda=490 N={terms * H}", timeit(lambda: th.mm(x16, Wcat, out_dtype=th.float32)),
          "max rel err", err, "(fp32 sgemm:", e32, ")")
    # padded pitch 496
    xp = th.zeros(rows, 496, dtype=th.float16, device=dev)
    xp[:, :K] = x16
    Wp = th.zeros(496, terms * H, dtype=th.float16, device=dev)
    Wp[:K] = Wcat
    print(f"terms {terms}: fp16 mm lda=496", timeit(lambda: th.mm(xp, Wp, out_dtype=th.float32)))
    xv = xp[:, :K]  # logical K = 490, pitch 496 (aligned rows)
    try:
        print(f"terms {terms}: fp16 mm K=490 pitch 496", timeit(lambda: th.mm(xv, Wcat, out_dtype=th.float32)))
    except Exception as e:
        print("view failed", e)
x3 = th.rand(4096, 101, 100, K, device=dev, dtype=th.float16)[:, 5]
print("strided obs slice contiguous?", x3.is_contiguous())
