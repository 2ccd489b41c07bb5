"""Debug: per-phase cycle stamps of the tcgen05 attention kernel (CTA 0, first work unit)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sam_road_b200 import _lib
lib = _lib.load()
B, s, heads, hd = 64, 32, 12, 64
D = heads * hd
def _time(fn, iters=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for win in (32, 14):
    g = torch.Generator().manual_seed(1)
    qkv = (torch.randn(B * s * s, 3 * D, generator=g)).to(torch.float16).cuda()
    bias = torch.zeros(3 * D).cuda()
    rel_h = (0.1 * torch.randn(2 * win - 1, hd, generator=g)).cuda()
    rel_w = (0.1 * torch.randn(2 * win - 1, hd, generator=g)).cuda()
    out = torch.empty(B * s * s, D, dtype=torch.float16, device="cuda")
    call = lambda: _lib.check(lib.samroad_op_attention(qkv.data_ptr(), bias.data_ptr(), rel_h.data_ptr(), rel_w.data_ptr(),
                                                       B, s, win, heads, hd, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "att")
    for rep in range(2):
        for label, mode in (("default (turn-taking)", 0), ("free-running + stagger", 12), ("free-running", 8)):
            lib.samroad_debug_force_simt_attention(mode)
            print(f"win={win} {label}: {_time(call):.1f} us")
    lib.samroad_debug_force_simt_attention(0)
    tr = torch.zeros(256, dtype=torch.int64, device="cuda")
    lib.samroad_debug_attention_trace(tr.data_ptr())
    for _ in range(2):
        _lib.check(lib.samroad_op_attention(qkv.data_ptr(), bias.data_ptr(), rel_h.data_ptr(), rel_w.data_ptr(),
                                            B, s, win, heads, hd, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "att")
    torch.cuda.synchronize()
    lib.samroad_debug_attention_trace(None)
    lib.samroad_debug_force_simt_attention(0)
    t = tr.cpu().tolist()
    nblk = 8 if win == 32 else 2
    t0 = t[0]
    print(f"win={win}: softmax warp phases per block (cycles): wait_S, ldtm, logits+max, exps, wait_pv, P stores, fence+arrive | cum")
    for jb in range(nblk):
        r = t[jb * 8: jb * 8 + 8]
        print(jb, [r[i + 1] - r[i] for i in range(7)], r[0] - t0, r[7] - t0)
    print(" unit starts (cycles since first):", [t[112 + i] - t[112] for i in range(8)])
    print(" unit epilogue starts - unit start:", [t[120 + i] - t[112 + i] for i in range(8)])
    print(" prologue (unit start -> first block):", [t[128 + i] - t[112 + i] for i in range(8)])
    print(" blocks (first block -> epilogue start):", [t[120 + i] - t[128 + i] for i in range(8)])
    print(" epilogue (O/l -> global):", [t[136 + i] - t[120 + i] for i in range(8)])
    print(" MMA S issue (g0,g1):", [(t[64 + jb * 2] - t0, t[65 + jb * 2] - t0) for jb in range(nblk)])
    print(" MMA PV issue (g0,g1):", [(t[96 + jb * 2] - t0, t[97 + jb * 2] - t0) for jb in range(nblk)])
