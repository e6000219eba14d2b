"""Phase timing of the tcgen05 attention kernels from in-kernel clock64 stamps (debug build only):
  CT_DEBUG_TIMING=1 python -m cleantransformer_b200.build
  CT_B200_LIB=cleantransformer_b200/libct_b200_dbg.so python tools/attn_timing.py
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cleantransformer_b200 import ops, _lib
from oracle import ct_oracle as O
B, H, S, D = 8, 16, 1024, 64
qkv = torch.randn(B, S, H, 3, D, device="cuda").bfloat16()
q, k, v = [qkv[..., i, :].permute(0, 2, 1, 3) for i in range(3)]
mask = torch.ones(B, S, dtype=torch.long, device="cuda")
kb2, fv = ops.attn_mask_prep(mask, H, 0, O.alibi_slopes(H).cuda())
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_longlong * 4096)()
for rep in range(2):
    o, lse2 = ops.attn_fwd(q, k, v, 0.125, True, -ops.FLT_MAX, kb2, fv)
    torch.cuda.synchronize()
    lib.ct_debug_timing(buf, 4096)
    f = list(buf)
    do = torch.randn_like(o); dq3 = torch.empty_like(qkv)
    dq, dk, dv = [dq3[..., i, :].permute(0, 2, 1, 3) for i in range(3)]
    ops.attn_bwd(do, q, k, v, o, lse2, dq, dk, dv, 0.125, True, -ops.FLT_MAX, kb2, fv)
    torch.cuda.synchronize()
    lib.ct_debug_timing(buf, 4096)
    b = list(buf)
print("FWD block 700 (thread 64 = first softmax warp): per kv tile: wait_s, pass1, pass2, wait_o(+PV)")
for j in range(8):
    t = f[2048 + 16 * j: 2048 + 16 * j + 5]
    if t[1] == 0: break
    nxt = f[2048 + 16 * (j + 1)]
    print(j, "wait_s", t[1] - t[0], "pass1", t[2] - t[1], "pass2", t[3] - t[2], "p_ready->o_full", t[4] - t[3], "o_update", (nxt - t[4]) if nxt else -1)
print("BWD block 700: per q tile: wait_sdp, compute, wait_dq, dq_red")
for it in range(8):
    t = b[16 * it: 16 * it + 6]
    if t[1] == 0: break
    print(it, "wait_sdp", t[1] - t[0], "compute", t[2] - t[1], "fence/arrive", t[3] - t[2], "wait_dq", t[4] - t[3], "dq_red", t[5] - t[4])

# ---- GEMM phase timing (block 5): epilogue thread and MMA thread, per tile ----
T, Hd = 8192, 1024
x = torch.randn(T, Hd, device="cuda").bfloat16()
for (N, K, name) in [(3072, 1024, "qkv"), (4096, 1024, "ffn1"), (1024, 4096, "ffn2")]:
    w = torch.randn(N, K, device="cuda").bfloat16(); a = torch.randn(T, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    for _ in range(2):
        ops.gemm(a, w, T, N, K, bias=bias)
    torch.cuda.synchronize()
    lib.ct_debug_timing_gemm(buf, 4096)
    g = list(buf)
    print("GEMM", name, "block 5: tile: epi_wait, epi_work | mma: wait_tempty, mainloop_issue")
    for i in range(6):
        e = g[8 * i: 8 * i + 3]; m = g[2048 + 8 * i: 2048 + 8 * i + 3]
        if e[1] == 0: break
        print(i, "epi_wait", e[1] - e[0], "epi_work", e[2] - e[1], "| mma_wait", m[1] - m[0], "mma_issue", m[2] - m[1],
              "tile_period", (g[2048 + 8 * (i + 1)] - m[0]) if g[2048 + 8 * (i + 1)] else -1)
