"""Back-to-back (asynchronous) searches: time per search with the launch latency pipelined away, next to
one search at a time and to an empty kernel.  usage: probe_pipeline.py <rows> [k]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import better_search_rag_rust_b200 as bsr

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2305
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
store = bsr.GpuVectorstore(None, True, dim=768, capacity_rows=n)
store.append_synthetic(0x5EED0001, 0, n)
eng = bsr.CudaShardEngine(store)
q = torch.from_numpy(store.get(0)).cuda().reshape(1, -1)
out = torch.empty((1, bsr.candidate_slots(k)), dtype=torch.int64, device="cuda")
x = torch.zeros(32, device="cuda")


def timed(fn, reps):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def one_at_a_time(fn, reps):
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(t))


search = lambda: eng.search_candidates(q, k, out=out)
empty = lambda: x.add_(1.0)
print(f"n={n} k={k}: pipelined {timed(search, 500):.1f} us/search, one at a time {one_at_a_time(search, 200):.1f} us; "
      f"empty kernel: pipelined {timed(empty, 500):.1f} us, one at a time {one_at_a_time(empty, 200):.1f} us", flush=True)
