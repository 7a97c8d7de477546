"""One 8-query exact pass over 1 M x 768 fp32 rows (ncu target for the QB = 8 scan kernel)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import better_search_rag_rust_b200 as bsr
from oracle import c_oracle as co
store = bsr.GpuVectorstore(None, True, dim=768, capacity_rows=1_000_000)
store.append_synthetic(0x5EED0001, 0, 1_000_000)
eng = bsr.CudaShardEngine(store)
q = torch.from_numpy(co.generate_rows(0x5EED1000, 0, 8, 768)).cuda()
out = torch.empty((8, bsr.candidate_slots(10)), dtype=torch.int64, device="cuda")
for _ in range(4):
    eng.search_candidates(q, 10, out=out)
torch.cuda.synchronize()
print("ok", store.last_scan_ms())
