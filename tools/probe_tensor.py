"""Batched-search timing probe (device-resident queries; not the bench contract).
usage: probe_tensor.py <rows> <f32|bf16> <k> <nq> [reps]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import better_search_rag_rust_b200 as bsr
from oracle import c_oracle as co

n = int(sys.argv[1]); dtype = sys.argv[2]; k = int(sys.argv[3]); nq = int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
store = bsr.GpuVectorstore(None, True, dim=768, capacity_rows=n, dtype=dtype)
store.append_synthetic(0x5EED0001, 0, n)
eng = bsr.CudaShardEngine(store)
q = torch.from_numpy(co.generate_rows(0x5EED1000, 0, nq, 768)).cuda()
out = torch.empty((nq, bsr.candidate_slots(k)), dtype=torch.int64, device="cuda")
for _ in range(2):
    eng.search_candidates(q, k, out=out)
torch.cuda.synchronize()
ms, scan = [], []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.search_candidates(q, k, out=out); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1)); scan.append(store.last_scan_ms())
m = float(np.median(ms))
print(f"n={n} {dtype} k={k} nq={nq}: batch ms median {m:.3f} (scan+finalize {np.median(scan):.3f}) -> {nq / m * 1e3:.0f} q/s, "
      f"{2.0 * n * 768 * nq / m / 1e9:.0f} TFLOP/s", flush=True)
