"""Quick device-side timing probe of the scan kernel (not the bench contract)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import better_search_rag_rust_b200 as bsr

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
store = bsr.GpuVectorstore(None, True, dim=768, capacity_rows=n)
t = time.time(); store.append_synthetic(0x5EED0001, 0, n); print("gen s", time.time() - t)
eng = bsr.CudaShardEngine(store)
q = torch.from_numpy(store.get(0)).cuda().reshape(1, -1)
for _ in range(3):
    eng.search_candidates(q, k)
torch.cuda.synchronize()
ms_scan, ms_all = [], []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = eng.search_candidates(q, k); e1.record(); torch.cuda.synchronize()
    ms_all.append(e0.elapsed_time(e1)); ms_scan.append(store.last_scan_ms())
b = store.scan_bytes()
print(f"n={n} k={k} bytes={b/1e9:.3f} GB")
print(f"scan  ms median {np.median(ms_scan):.4f} min {min(ms_scan):.4f} -> {b/np.median(ms_scan)/1e6:.1f} GB/s (best {b/min(ms_scan)/1e6:.1f})")
print(f"total ms median {np.median(ms_all):.4f} min {min(ms_all):.4f} -> {1e3/np.median(ms_all):.1f} q/s")
idx, dist, cnt = eng.unpack(c, k); print(idx[0][:5], dist[0][:5])
