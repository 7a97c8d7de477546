"""Query-batch throughput probe (exact path): queries/s vs batch size, device-resident queries."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import better_search_rag_rust_b200 as bsr
from oracle import c_oracle as co

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
store = bsr.GpuVectorstore(None, True, dim=768, capacity_rows=n)
store.append_synthetic(0x5EED0001, 0, n)
eng = bsr.CudaShardEngine(store)
base = None
for nq in (1, 2, 4, 8, 16, 64, 256):
    qs = co.generate_rows(0x5EED1000, 0, nq, 768)
    qs[0] = store.get(0)
    q = torch.from_numpy(qs).cuda()
    out = torch.empty((nq, bsr.candidate_slots(k)), dtype=torch.int64, device="cuda")
    for _ in range(3):
        eng.search_candidates(q, k, out=out)
    torch.cuda.synchronize()
    reps = max(3, 200 // nq)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.search_candidates(q, k, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    qps = nq / ms * 1e3
    base = base or qps
    print(f"nq={nq:4d} ms/batch={ms:8.3f} q/s={qps:10.1f} x{qps/base:5.2f}  eff GB/s per query-pass={store.scan_bytes()*nq/ms/1e6:9.1f}")
# correctness of a 13-query batch against the oracle on a small store
rows = co.generate_rows(7, 0, 5000, 768)
st2 = bsr.GpuVectorstore(None, True); st2.append_many(rows)
e2 = bsr.CudaShardEngine(st2)
qs = np.concatenate([rows[:3], co.generate_rows(8, 0, 10, 768)])
idx, dist, cnt = e2.local_top_k(qs, k)
for i, qq in enumerate(qs):
    wi, wd = co.compute_local_top_k(rows, 0, 1, k, qq)
    assert idx[i].tolist() == wi.tolist() and dist[i].view(np.uint32).tolist() == wd.view(np.uint32).tolist(), i
print("batch parity ok")
