"""Debug: per-phase cycle stamps of the tcgen05 attention kernel (CTA 0, first work unit)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sam_road_b200 import _lib
lib = _lib.load()
B, s, heads, hd = 64, 32, 12, 64
D = heads * hd
for win in (32, 14):
    g = torch.Generator().manual_seed(1)
    qkv = (torch.randn(B * s * s, 3 * D, generator=g)).to(torch.float16).cuda()
    bias = torch.zeros(3 * D).cuda()
    rel_h = (0.1 * torch.randn(2 * win - 1, hd, generator=g)).cuda()
    rel_w = (0.1 * torch.randn(2 * win - 1, hd, generator=g)).cuda()
    out = torch.empty(B * s * s, D, dtype=torch.float16, device="cuda")
    tr = torch.zeros(128, dtype=torch.int64, device="cuda")
    lib.samroad_debug_attention_trace(tr.data_ptr())
    for _ in range(2):
        _lib.check(lib.samroad_op_attention(qkv.data_ptr(), bias.data_ptr(), rel_h.data_ptr(), rel_w.data_ptr(),
                                            B, s, win, heads, hd, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "att")
    torch.cuda.synchronize()
    lib.samroad_debug_attention_trace(None)
    t = tr.cpu().tolist()
    nblk = 8 if win == 32 else 2
    t0 = t[0]
    print(f"win={win}: softmax warp phases per block (cycles): wait_S, ldtm, logits+max, exps, wait_pv, P stores, fence+arrive | cum")
    for jb in range(nblk):
        r = t[jb * 8: jb * 8 + 8]
        print(jb, [r[i + 1] - r[i] for i in range(7)], r[0] - t0, r[7] - t0)
    print(" MMA S issue (g0,g1):", [(t[64 + jb * 2] - t0, t[65 + jb * 2] - t0) for jb in range(nblk)])
    print(" MMA PV issue (g0,g1):", [(t[96 + jb * 2] - t0, t[97 + jb * 2] - t0) for jb in range(nblk)])
