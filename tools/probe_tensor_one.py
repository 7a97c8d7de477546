import sys
import numpy as np, torch
sys.path.insert(0, ".")
import better_search_rag_rust_b200 as bsr
from oracle import c_oracle as co
n = int(sys.argv[1]); k = int(sys.argv[2]); nq = int(sys.argv[3])
store = bsr.GpuVectorstore(None, True, dim=768, capacity_rows=n, dtype="bf16")
store.append_synthetic(0x5EED0001, 0, n)
eng = bsr.CudaShardEngine(store)
q = torch.from_numpy(co.generate_rows(0x5EED1000, 0, nq, 768)).cuda()
out = torch.empty((nq, bsr.candidate_slots(k)), dtype=torch.int64, device="cuda")
for _ in range(3):
    eng.search_candidates(q, k, out=out)
torch.cuda.synchronize()
print("ok", store.last_scan_ms())
