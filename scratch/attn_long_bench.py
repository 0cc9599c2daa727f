import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
Hq, Hkv, dh, NL, PAGE = 8, 1, 256, 18, 64
W = (Hq + 2) * dh
def rnd(*s): return (torch.randn(*s, device="cuda") * 0.02).bfloat16()
inv_freq=(1.0/(10000.0**(torch.arange(0,dh,2,dtype=torch.int64).float()/dh))).cuda()
for B, kvlen in ((64, 324), (32, 1092), (8, 4164), (1, 324), (8, 324)):
    max_pages = (kvlen + 63) // 64 + 1
    k_pages=[rnd(B*max_pages, PAGE, dh) for _ in range(NL)]; v_pages=[rnd(B*max_pages, PAGE, dh) for _ in range(NL)]
    table=torch.arange(B*max_pages, device="cuda", dtype=torch.int32).view(B,max_pages).contiguous()
    kvl=torch.full((B,), kvlen, device="cuda", dtype=torch.int32); posd=kvl.clone()
    qkvf=torch.randn(B, W, device="cuda")*0.5
    out=torch.empty(B, Hq*dh, device="cuda", dtype=torch.bfloat16)
    def attn(i):
        _lib.check(L.pg_attention_decode_fused(qkvf.data_ptr(), posd.data_ptr(), kvl.data_ptr(), inv_freq.data_ptr(), k_pages[i].data_ptr(), v_pages[i].data_ptr(), table.data_ptr(), out.data_ptr(), B, Hq, Hkv, dh, PAGE, B*max_pages, max_pages, 1.0/16, _lib.stream()), "attn")
    for i in range(NL): attn(i)
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(NL): attn(i)
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * NL)
    nbytes = B * kvlen * dh * 4
    print(f"B={B:3d} kv={kvlen:5d}: {us:7.2f} us  {nbytes / us / 1e3:7.1f} GB/s  ideal {nbytes / 6550.7e3:6.2f} us")
    del k_pages, v_pages
