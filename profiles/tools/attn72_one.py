"""One shape of the dh = 72 tcgen05 prefill attention (SigLIP 896 px: 8 images x 16 heads x 4096 tokens), for ncu."""
import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B, H, dh = 32768 // N, 16, 72
D = H * dh
qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.7).bfloat16()
out = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    _lib.check(L.pg_attention_prefill(qkv.data_ptr(), qkv.data_ptr() + 2 * D, qkv.data_ptr() + 4 * D, out.data_ptr(), B, H, N, N, dh, 1,
               N * 3 * D, 3 * D, 0, dh, N * 3 * D, 3 * D, dh, N * D, D, 0, dh, dh ** -0.5, _lib.stream()), "a")
torch.cuda.synchronize()
print("ok")
