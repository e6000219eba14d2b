"""Phase timing of the v2 attention kernels from in-kernel clock64 stamps (debug build only):
  CT_DEBUG_TIMING=1 python -m cleantransformer_b200.build
  CT_B200_LIB=cleantransformer_b200/libct_b200_dbg.so python tools/attn_timing2.py
Stamped thread: block 700, thread 64 (first softmax / compute warp, lane 0)."""
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
names_f = ["wait_s_full", "kb_stage+bar", "tmem_ld(4)+wait", "scale_max+s_free", "wait_o_full(j-1)", "exp+store", "rescale_O", "arrive_p_ready"]
print("FWD v2, block 700: cycles per phase per kv tile")
for j in range(8):
    t = f[2048 + 16 * j: 2048 + 16 * j + 9]
    if t[1] == 0: break
    print(j, {n: t[i + 1] - t[i] for i, n in enumerate(names_f) if t[i + 1] and t[i]}, "tile_total", t[8] - t[0])
names_b = ["wait_sdp_full", "issue_ld+wait_mma_done", "stage+tmem_ld_wait", "arrive+bar_sync", "chunk0", "dq_drain", "chunk1", "fence+arrive"]
print("BWD v2, block 700: cycles per phase per q tile")
for it in range(8):
    t = b[16 * it: 16 * it + 9]
    if t[1] == 0: break
    if b[16 * it + 9] and t[5]:
        print("   v3: wait for mma_done(it-1) inside dq_drain:", b[16 * it + 9] - t[5])
    print(it, {n: t[i + 1] - t[i] for i, n in enumerate(names_b) if t[i + 1] and t[i]}, "tile_total", t[8] - t[0],
          "next_start_gap", (b[16 * (it + 1)] - t[8]) if b[16 * (it + 1)] else None)

print("BWD whole-CTA timeline of every 64th block (cycles): setup, loop, wait_dkv, epilogue, exit_sync | total; globaltimer ns: start offset, duration")
g0 = min(b[1024 + 8 * k + 6] for k in range(16) if b[1024 + 8 * k + 6])
for blk in range(16):
    t = b[1024 + 8 * blk: 1024 + 8 * blk + 8]
    if not t[0]:
        continue
    bid = blk * 64 + 60
    print(bid, "tiles", 8 - bid % 8, [t[i + 1] - t[i] for i in range(5)], "total", t[5] - t[0], "| start_ns", t[6] - g0, "dur_ns", t[7] - t[6])
