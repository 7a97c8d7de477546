"""Throughput probe of the tcgen05 batched path (BF16 store)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import better_search_rag_rust_b200 as bsr
from oracle import c_oracle as co

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
store = bsr.GpuVectorstore(None, True, dim=768, capacity_rows=n, dtype="bf16")
store.append_synthetic(0x5EED0001, 0, n)
eng = bsr.CudaShardEngine(store)
for nq in (1, 8, 128, 256, 1024):
    qs = co.generate_rows(0x5EED1000, 0, nq, 768)
    q = torch.from_numpy(qs).cuda()
    out = torch.empty((nq, bsr.candidate_slots(k)), dtype=torch.int64, device="cuda")
    for _ in range(2):
        eng.search_candidates(q, k, out=out)
    torch.cuda.synchronize()
    reps = 3
    store.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.search_candidates(q, k, out=out)
    e1.record(); torch.cuda.synchronize()
    scan_ms, n_scans = store.scan_profile(); store.set_profiling(False)
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * n * 768 * nq
    print(f"nq={nq:5d} ms/batch={ms:9.3f} q/s={nq/ms*1e3:10.1f} scan-kernels ms={scan_ms/reps:9.3f} ({n_scans//reps} launches) "
          f"TFLOP/s={flops/ms/1e9:8.1f} store GB/s per pass={store.scan_bytes()*max(1,(nq+127)//128)/ (scan_ms/reps)/1e6 if nq>=32 else store.scan_bytes()*nq/(scan_ms/reps)/1e6:9.1f}")
