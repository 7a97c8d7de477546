"""Device-side timing probe of the exact single-query scan for either store dtype (not the bench contract).
usage: probe_dtype.py <rows> <f32|bf16> [k] [reps]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import better_search_rag_rust_b200 as bsr

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 30
store = bsr.GpuVectorstore(None, True, dim=768, capacity_rows=n, dtype=dtype)
store.append_synthetic(0x5EED0001, 0, n)
store.set_tensor_min_batch(0)
eng = bsr.CudaShardEngine(store)
q = torch.from_numpy(store.get(0)).cuda().reshape(1, -1)
for _ in range(5):
    eng.search_candidates(q, k)
torch.cuda.synchronize()
ms = []
for _ in range(reps):
    eng.search_candidates(q, k); torch.cuda.synchronize(); ms.append(store.last_scan_ms())
b = store.scan_bytes()
print(f"n={n} {dtype} k={k} bytes={b/1e9:.3f} GB scan ms median {np.median(ms):.4f} min {min(ms):.4f} -> "
      f"{b/np.median(ms)/1e6:.1f} GB/s (best {b/min(ms)/1e6:.1f})", flush=True)
