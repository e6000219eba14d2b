"""Multi-GPU check of csrc/comm.cu + ddp.py. Launch with torchrun (one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/ddp_check.py

(1) ct_allreduce_bucket / ct_broadcast vs torch.distributed (NCCL) on random data, several sizes,
    both modes;  (2) all-reduce bandwidth on 25 MiB / 256 MiB / 1 GiB ranges vs NCCL;
(3) DDP wrapper on a small Bloom: gradients after backward == mean over ranks of the local
    gradients, p2p path == nccl path;  writes gpurun_out/ddp_check_rank0.json.
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from cleantransformer_b200 import _lib
    from cleantransformer_b200.ddp import DistributedDataParallel, _CudaView
    from cleantransformer_b200.models import modeling_bloom as mb
    lib = _lib.load()
    res = {"world": world}

    # ---- raw collective -------------------------------------------------------------------
    n_total = (1 << 28) + (1 << 26)  # 1.25 GiB of f32: 1 GiB range + one-shot staging room
    local_ptr = ctypes.c_void_p(); dh = ctypes.create_string_buffer(64); sh = ctypes.create_string_buffer(64)
    _lib.check(lib.ct_comm_init(rank, world, local, n_total * 4, ctypes.byref(local_ptr), dh, sh), "ct_comm_init")
    gathered = [None] * world
    dist.all_gather_object(gathered, (dh.raw, sh.raw))
    _lib.check(lib.ct_comm_connect(b"".join(g[0] for g in gathered), b"".join(g[1] for g in gathered)), "connect")
    dist.barrier()
    buf = torch.as_tensor(_CudaView(local_ptr.value, n_total), device=dev)
    st = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(1234 + rank)
    errs = {}
    for (off, cnt, mode) in [(0, 4, 0), (64, 1000, 0), (4096, 1 << 20, 0), (128, 12345 * 4, 0), (0, 1 << 16, 1), (256, 1024, 1)]:
        x = torch.randn(cnt, device=dev)
        buf[off:off + cnt].copy_(x)
        ref = x.clone(); dist.all_reduce(ref); ref /= world
        torch.cuda.synchronize(); dist.barrier()
        _lib.check(lib.ct_allreduce_bucket(off, cnt, 1.0 / world, mode, 0, st), "allreduce")
        torch.cuda.synchronize()
        errs["ar_%d_%d_m%d" % (off, cnt, mode)] = rel(buf[off:off + cnt], ref)
    x = torch.randn(1 << 18, device=dev); buf[0:1 << 18].copy_(x)
    ref = x.clone(); dist.broadcast(ref, 1 % world)
    torch.cuda.synchronize(); dist.barrier()
    _lib.check(lib.ct_broadcast(0, 1 << 18, 1 % world, st), "broadcast")
    torch.cuda.synchronize()
    errs["bcast"] = rel(buf[0:1 << 18], ref)
    res["collective_errors"] = errs

    bw = {}
    for mib in (25, 256, 1024):
        cnt = mib * (1 << 20) // 4
        buf[:cnt].normal_()
        t = buf[:cnt].clone()
        for name in ("p2p", "nccl"):
            for ctas in ((16, 32, 64) if name == "p2p" else (0,)):
                torch.cuda.synchronize(); dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 5
                for it in range(reps + 2):
                    if it == 2:
                        e0.record()
                    if name == "p2p":
                        _lib.check(lib.ct_allreduce_bucket(0, cnt, 1.0 / world, 0, ctas, st), "allreduce")
                    else:
                        dist.all_reduce(t)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                tt = torch.tensor([ms], device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                key = "%s%s_%dMiB" % (name, ("_c%d" % ctas) if ctas else "", mib)
                # bus bandwidth convention: 2(W-1)/W * bytes / time
                bw[key] = {"ms": float(tt), "busbw_GBs": 2 * (world - 1) / world * cnt * 4 / (float(tt) * 1e-3) / 1e9}
    res["bandwidth"] = bw
    _lib.check(lib.ct_comm_finalize(), "finalize")
    dist.barrier()

    # ---- DDP wrapper end to end -----------------------------------------------------------------
    cfg = dict(vocab_size=4096, hidden_size=256, n_layer=2, num_attention_heads=4)

    def build(seed):
        torch.manual_seed(seed)
        with torch.device(dev):
            m = mb.BloomForCausalLM(mb.BloomConfig(**cfg))
        with torch.no_grad():
            for _, p in m.named_parameters():
                if p.dim() >= 2:
                    p.normal_(0, 0.02)
        m._tie_weight(); m.train()
        return m

    g = torch.Generator().manual_seed(77 + rank)
    ids = torch.randint(3, 4096, (4, 256), generator=g).to(dev)
    mask = torch.ones(4, 256, dtype=torch.long, device=dev)
    base = build(5)  # same init on every rank; local gradients
    (l, _, _), _ = base(input_ids=ids, attention_mask=mask, labels=ids); l.backward()
    local_grads = {n: p.grad.detach().clone() for n, p in base.named_parameters()}
    mean_grads = {}
    for n, gr in local_grads.items():
        t = gr.clone(); dist.all_reduce(t); mean_grads[n] = t / world
    out = {}
    # "ce" (copy-engine transport) was written after the round's GPU budget was spent: opt-in until it has run once
    for comm in ("p2p", "nccl") + (("ce",) if os.environ.get("CT_TEST_EXPERIMENTAL") else ()):
        m = build(5 + rank * (comm == "p2p"))  # p2p run starts from rank-dependent weights: ctor must sync
        ddp = DistributedDataParallel(m, device_ids=[local], comm=comm, bucket_cap_mb=1)
        if comm == "p2p":
            w0 = m.bloom.blocks[0].mlp.dense_h_to_4h.weight.detach().clone()
            wr = w0.clone(); dist.broadcast(wr, 0)
            out["param_sync"] = rel(w0, wr)
        for step in range(2):
            for p in m.parameters():
                p.grad = None
            (l, _, _), _ = ddp(input_ids=ids, attention_mask=mask, labels=ids)
            l.backward()
        torch.cuda.synchronize()
        # (p2p: weights differ from `base` on rank>0 before the constructor's sync; after it all ranks == seed 5)
        worst = 0.0
        for n, p in m.named_parameters():
            worst = max(worst, rel(p.grad, mean_grads[n]))
        out[comm + "_vs_mean"] = worst
        out["buckets_" + comm] = len(ddp.buckets)
        if comm in ("p2p", "ce"):
            _lib.check(lib.ct_comm_finalize(), "finalize")
        dist.barrier()
    res["ddp"] = out
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ddp_check_rank0.json"), "w"), indent=1)
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
