import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
V = 257216
bad = 0
junk = [torch.randn(1 << 20, device="cuda") for _ in range(8)]
for it in range(300):
    logits = torch.zeros(2, V, device="cuda")
    logits[0, 108] = 50.0
    logits[1, 5] = 30.0
    out = torch.full((2,), -9, device="cuda", dtype=torch.int32)
    cnt = torch.full((2,), -7, device="cuda", dtype=torch.int32)
    if it % 3 == 0:
        junk[it % 8].normal_()   # unrelated work just before
    _lib.check(L.pg_sample_top_p(logits.data_ptr(), V, out.data_ptr(), cnt.data_ptr(), 2, V, 1.25, 0.9, 1, 0, _lib.stream()), "top-p")
    torch.cuda.synchronize()
    if out.tolist() != [108, 5] or cnt.tolist() != [1, 1]:
        bad += 1
        if bad <= 10: print("iter", it, out.tolist(), cnt.tolist())
print("bad", bad, "of 300")
