"""End-to-end latency of small stores through bsr_search (host buffers), plus the scan-kernel device time.
usage: probe_small.py <rows> [k] [reps]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import better_search_rag_rust_b200 as bsr

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2305
k = int(sys.argv[2]) if len(sys.argv) > 2 else 50
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
store = bsr.GpuVectorstore(None, True, dim=768, capacity_rows=n)
store.append_synthetic(0x5EED0001, 0, n)
eng = bsr.CudaShardEngine(store)
q = store.get(0).reshape(1, -1)
import os
if os.environ.get("PROBE_QUERY") == "random":   # not a store row: no lane ever runs the identity compare
    q = np.random.default_rng(7).uniform(-1, 1, size=(1, 768)).astype(np.float32)
from better_search_rag_rust_b200 import _lib
import ctypes as C
L = _lib.lib()
idx = np.zeros((1, k), np.uint64); dist = np.zeros((1, k), np.float32); cnt = np.zeros(1, np.uint32)
args = (store.handle, q.ctypes.data, 1, k, idx.ctypes.data, dist.ctypes.data, cnt.ctypes.data)
for _ in range(50):
    L.bsr_search(*args)
t = []
for _ in range(reps):
    t0 = time.perf_counter_ns(); L.bsr_search(*args); t.append(time.perf_counter_ns() - t0)
t = np.array(t) / 1e3
ms = []
for _ in range(20):
    L.bsr_search(*args)
    try:
        ms.append(store.last_scan_ms() * 1e3)
    except Exception:
        ms.append(float("nan"))
print(f"n={n} k={k}: e2e us median {np.median(t):.1f} p10 {np.percentile(t,10):.1f} min {t.min():.1f}; scan kernel us median {np.median(ms):.1f}", flush=True)
