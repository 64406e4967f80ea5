"""Feasibility: do two half-batches on two streams overlap the agent GEMMs of one with the env kernel of the other?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import bench
w = dict(bench.WORKLOADS["c3"]); T = 100
def make(B, seed, prio=0):
    ww = dict(w); ww["B"] = B
    g = th.Generator(device="cuda").manual_seed(seed)
    planes = th.rand(B, T, ww["n"], ww["m"], device="cuda", generator=g)
    runner, buffer, _ = bench.build_runner(ww, seed, planes)
    return runner, buffer
def episode_serial(r):
    with th.no_grad():
        r.reset(); r._rollout_loop(False)
def episode_dual(ra, rb, sa, sb):
    with th.no_grad():
        with th.cuda.stream(sa): ra.reset(); ra.mac.init_hidden(ra.batch_size)
        with th.cuda.stream(sb): rb.reset(); rb.mac.init_hidden(rb.batch_size)
        for t in range(T):
            for r, s in ((ra, sa), (rb, sb)):
                with th.cuda.stream(s):
                    a = r.mac.select_actions(r.batch, t_ep=t, t_env=0, test_mode=False)
                    r.env.step(a, r.batch); r.batch.agent_in_t = t + 1
def timeit(fn, n=3):
    fn(); th.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    th.cuda.synchronize(); return (time.perf_counter() - t0) / n
full, _ = make(4096, 1)
t_full = timeit(lambda: episode_serial(full))
print("one batch of 4096:      %.1f ms/episode" % (t_full * 1e3))
del full; th.cuda.empty_cache()
ra, _ = make(2048, 2); rb, _ = make(2048, 3)
t_ser = timeit(lambda: (episode_serial(ra), episode_serial(rb)))
print("two halves, one stream: %.1f ms/episode" % (t_ser * 1e3))
for pri in (False, True):
    sa = th.cuda.Stream(priority=0); sb = th.cuda.Stream(priority=-1 if pri else 0)
    t_dual = timeit(lambda: episode_dual(ra, rb, sa, sb))
    print("two halves, two streams%s: %.1f ms/episode" % (" (prio)" if pri else "", t_dual * 1e3))
