"""Multi-GPU parity + bandwidth check of csrc/comm.cu + ddp.py. Launch with torchrun (one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 tools/ddp_check.py [--quick] [--out gpurun_out/ddp_check_wN.json]

tests/test_gpu_multi.py spawns exactly this (--quick) for world 2 / 4 / 8 when that many GPUs are visible.

(1) ct_allreduce_bucket in every available mode (two-shot unicast, one-shot, NVLS multimem) and ct_broadcast against
    torch.distributed over NCCL: BIT-IDENTICAL on integer-valued data (every partial sum exact, so the summation order
    cannot matter), <= 1e-6 on random data, and bit-identical ACROSS ranks (replicas must not drift);
(2) (not --quick) all-reduce bus bandwidth on 25 MiB / 256 MiB / 1 GiB vs NCCL;
(3) the DDP wrapper on a small Bloom (tied table: early dense + sparse token-row exchange): gradients after backward
    == mean over ranks of the local gradients, for comm='p2p' and comm='nccl'; == torch DDP over NCCL around the
    ORACLE (the reference's arithmetic) in fp32 within the bf16 bound, with the oracle under bf16 autocast beside it;
    dense buckets bit-identical across ranks; parameter sync from rank 0;
(4) GPT with segment_ids (three writes to the tied table, modeling_gpt.py:186-188);
(5) the whole DDP step replayed from a CUDA graph (device-side epochs) == the eager DDP step.
Exit code != 0 on any violated bound; rank 0 writes the JSON.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

FAILS = []


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


def check(cond, what):
    if not cond:
        FAILS.append(what)


def same_on_all_ranks(t):
    """Is this tensor bit-identical on every rank?"""
    bits = t.detach().contiguous().view(-1).view(torch.int32).to(torch.int64)
    s = torch.stack([bits.sum(), (bits * torch.arange(1, bits.numel() + 1, device=t.device) % 1000003).sum()])
    lo, hi = s.clone(), s.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool((lo == hi).all())


class OneParam(torch.nn.Module):
    def __init__(self, n, dev):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(n, device=dev))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from cleantransformer_b200 import _lib
    from cleantransformer_b200.ddp import DistributedDataParallel
    from cleantransformer_b200.graphs import GraphedTrainStep
    from cleantransformer_b200.models import modeling_bloom as mb, modeling_gpt as mg
    from oracle import ct_oracle as O
    lib = _lib.load()
    res = {"world": world}

    # ---- raw collectives on the wrapper's own symmetric buffer ---------------------------------------------
    n_total = ((1 << 28) + (1 << 26)) if not args.quick else (1 << 24)
    holder = DistributedDataParallel(OneParam(n_total, dev), device_ids=[local], comm="p2p")
    buf = holder.arena.grad
    res["nvls"] = bool(holder.nvls)
    flags = (ctypes.c_int * 2)()
    lib.ct_comm_info(flags)
    res["vmm"] = bool(flags[0])
    st = torch.cuda.current_stream().cuda_stream
    modes = [(3, "unicast2shot"), (1, "oneshot")] + ([(2, "nvls")] if holder.nvls else [])
    errs = {}
    torch.manual_seed(1234 + rank)
    for mode, mname in modes:
        for (off, cnt) in [(0, 4), (64, 1000), (4096, 1 << 20), (128, 12345 * 4), (0, 1 << 16), (256, 1024)]:
            if mode == 1 and cnt > (1 << 16):
                continue
            for kind in ("int", "randn"):
                x = torch.randint(-8, 9, (cnt,), device=dev).float() if kind == "int" else torch.randn(cnt, device=dev)
                buf[off:off + cnt].copy_(x)
                ref = x.clone(); dist.all_reduce(ref); ref /= world
                torch.cuda.synchronize(); dist.barrier()
                _lib.check(lib.ct_allreduce_bucket(off, cnt, 1.0 / world, mode, 0, st), "allreduce")
                torch.cuda.synchronize()
                got = buf[off:off + cnt]
                key = "%s_%d_%d_%s" % (mname, off, cnt, kind)
                if kind == "int" and (world & (world - 1)) == 0:
                    errs[key] = 0.0 if torch.equal(got, ref) else rel(got, ref) + 1e-30
                    check(torch.equal(got, ref), key + ": not bit-identical to NCCL")
                else:
                    errs[key] = rel(got, ref)
                    check(errs[key] <= 1e-6, key + ": %g" % errs[key])
                check(same_on_all_ranks(got), key + ": ranks differ")
    x = torch.randn(1 << 18, device=dev); buf[0:1 << 18].copy_(x)
    ref = x.clone(); dist.broadcast(ref, 1 % world)
    torch.cuda.synchronize(); dist.barrier()
    _lib.check(lib.ct_broadcast(0, 1 << 18, 1 % world, st), "broadcast")
    _lib.check(lib.ct_comm_barrier(st), "barrier")
    torch.cuda.synchronize()
    check(torch.equal(buf[0:1 << 18], ref), "broadcast: not bit-identical to NCCL")
    errs["bcast"] = rel(buf[0:1 << 18], ref)
    res["collective_errors"] = errs

    if not args.quick:
        bw = {}
        for mib in (25, 256, 1024):
            cnt = mib * (1 << 20) // 4
            buf[:cnt].normal_()
            t = buf[:cnt].clone()
            variants = [("unicast", 3, c) for c in (16, 32, 64)] + \
                       ([("nvls", 2, c) for c in (8, 16, 32, 64)] if holder.nvls else []) + [("nccl", -1, 0)]
            for name, mode, ctas in variants:
                torch.cuda.synchronize(); dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 5
                for it in range(reps + 2):
                    if it == 2:
                        e0.record()
                    if mode >= 0:
                        _lib.check(lib.ct_allreduce_bucket(0, cnt, 1.0 / world, mode, ctas, st), "allreduce")
                    else:
                        dist.all_reduce(t)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                tt = torch.tensor([ms], device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                key = "%s%s_%dMiB" % (name, ("_c%d" % ctas) if ctas else "", mib)
                # bus bandwidth convention: 2(W-1)/W * bytes / time
                bw[key] = {"ms": float(tt), "busbw_GBs": 2 * (world - 1) / world * cnt * 4 / (float(tt) * 1e-3) / 1e9}
        res["bandwidth"] = bw
    holder.close()
    del holder, buf

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
    sd0 = {k: v.detach().clone() for k, v in base.state_dict().items() if k != "lm_head.weight"}

    # the reference's way: torch DDP over NCCL around the oracle (fp32 and bf16 autocast)
    class OracleNet(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.keys = list(sd0.keys())
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(sd0[k].clone()) for k in self.keys])

        def forward(self, ids, mask):
            sd = dict(zip(self.keys, self.ps))
            (loss, _, _), _ = O.bloom_causal_lm(ids, mask, sd, 2, 4, 1e-5, labels=ids, training=True)
            return loss

    ref_grads = {}
    for tag, ac in (("fp32", False), ("autocast", True)):
        net = OracleNet()
        dd = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local])
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            loss = dd(ids, mask)
        loss.backward()
        ref_grads[tag] = {k: p.grad.detach().clone() for k, p in zip(net.keys, net.ps)}
        del dd, net
    out = {}
    for comm in ("p2p", "nccl"):
        m = build(5 + rank * (comm == "p2p"))  # p2p run starts from rank-dependent weights: ctor must sync
        ddp = DistributedDataParallel(m, device_ids=[local], comm=comm, bucket_cap_mb=1)
        if comm == "p2p":
            out["nvls"] = bool(ddp.nvls)
            w0 = m.bloom.blocks[0].mlp.dense_h_to_4h.weight.detach().clone()
            wr = w0.clone(); dist.broadcast(wr, 0)
            out["param_sync"] = rel(w0, wr)
            check(torch.equal(w0, wr), "parameter sync from rank 0")
        for step in range(2):
            for p in m.parameters():
                p.grad = None
            (l, _, _), _ = ddp(input_ids=ids, attention_mask=mask, labels=ids)
            l.backward()
        torch.cuda.synchronize()
        worst = worst_ref = worst_ac = 0.0
        for n, p in m.named_parameters():
            key = "bloom.word_embeddings.weight" if n == "lm_head.weight" else n
            worst = max(worst, rel(p.grad, mean_grads[n]))
            e_o, e_a = rel(p.grad, ref_grads["fp32"][key]), rel(ref_grads["autocast"][key], ref_grads["fp32"][key])
            worst_ref, worst_ac = max(worst_ref, e_o), max(worst_ac, e_a)
            check(e_o <= max(1.5 * e_a, 8e-3), "%s grad %s vs torch-DDP(oracle fp32): %g (autocast oracle %g)" % (comm, n, e_o, e_a))
            if comm == "p2p" and "word_embeddings.weight" not in n and n != "lm_head.weight":
                check(same_on_all_ranks(p.grad), "p2p grad %s differs across ranks" % n)
        # local gradients carry split-K / dQ atomics noise that bf16 re-rounding amplifies: 4e-3, not 1e-6
        check(worst <= 4e-3, "%s gradients vs mean of local gradients: %g" % (comm, worst))
        out[comm + "_vs_mean"] = worst
        out[comm + "_vs_torchddp_oracle_fp32"] = worst_ref
        out["oracle_autocast_vs_fp32"] = worst_ac
        out["buckets_" + comm] = len(ddp.buckets)
        if comm == "p2p":
            # ---- the same step replayed from a CUDA graph (collectives included) ----
            eager = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
            for p in m.parameters():
                p.grad = None
            gstep = GraphedTrainStep(ddp, dict(input_ids=ids, attention_mask=mask, labels=ids))
            worst_g = 0.0
            for it in range(2):
                loss_g = gstep(input_ids=ids, attention_mask=mask, labels=ids)
                torch.cuda.synchronize()
                for n, p in m.named_parameters():
                    worst_g = max(worst_g, rel(p.grad, eager[n]))
            check(abs(float(loss_g) - float(l)) <= 1e-5 * abs(float(l)), "graphed DDP loss")
            check(worst_g <= 4e-3, "graphed DDP gradients vs eager DDP: %g" % worst_g)
            out["graph_vs_eager"] = worst_g
            del gstep
        ddp.close()
        dist.barrier()
    # ---- the sparse tied-table exchange in several pieces (staging area smaller than the batch) ----
    os.environ["CT_DDP_STAGE_TOKENS"] = "300"   # 1024 tokens per rank -> 4 exchanges
    try:
        m = build(5)
        ddp = DistributedDataParallel(m, device_ids=[local], comm="p2p", bucket_cap_mb=1)
        (l, _, _), _ = ddp(input_ids=ids, attention_mask=mask, labels=ids)
        l.backward()
        torch.cuda.synchronize()
        worst = max(rel(p.grad, mean_grads[n]) for n, p in m.named_parameters())
        check(worst <= 4e-3, "chunked sparse exchange vs mean of local gradients: %g" % worst)
        out["chunked_sparse_exchange_vs_mean"] = worst
        ddp.close()
    finally:
        del os.environ["CT_DDP_STAGE_TOKENS"]
    dist.barrier()
    res["ddp"] = out

    # ---- GPT with segment_ids: the tied table receives three gradient writes ----
    gcfg = dict(vocab_size=2048, n_embd=256, n_positions=256, n_layer=2, n_head=4, n_ctx=256, embd_pdrop=0.0,
                attn_pdrop=0.0, resid_pdrop=0.0)

    def gbuild():
        torch.manual_seed(9)
        with torch.device(dev):
            mm = mg.GPTLMHeadModel(mg.GPTConfig(**gcfg), version="gpt2")
        mm._tie_weights()
        return mm.eval()

    gg = torch.Generator().manual_seed(99 + rank)
    gids = torch.randint(1, 2048, (2, 128), generator=gg).to(dev)
    gseg = torch.randint(1, 2048, (2, 128), generator=gg).to(dev)
    gmask = torch.ones_like(gids)

    def gloss(model):
        (logits, _), _ = model(gids, attention_mask=gmask, segment_ids=gseg)
        return torch.nn.functional.cross_entropy(logits.float().view(-1, 2048), gids.view(-1))

    gb = gbuild(); gloss(gb).backward()
    gmean = {}
    for n, p in gb.named_parameters():
        t = p.grad.detach().clone(); dist.all_reduce(t); gmean[n] = t / world
    gm = gbuild()
    gddp = DistributedDataParallel(gm, device_ids=[local], comm="p2p", bucket_cap_mb=1)
    gloss(gddp).backward()
    torch.cuda.synchronize()
    worst = max(rel(p.grad, gmean[n]) for n, p in gm.named_parameters())
    check(worst <= 4e-3, "GPT + segment_ids through DDP: %g" % worst)
    res["gpt_segment_ids_vs_mean"] = worst
    gddp.close()

    # ---- BERT (config 5: separate q/k/v Linears, post-LN, dropout 0.1 in train mode) through the DDP wrapper ----
    from cleantransformer_b200 import functional as F
    from cleantransformer_b200.models import modeling_bert as mbert
    bcfg = mbert.BertConfig(num_hidden_layers=2, num_labels=28)  # hidden / attention dropout 0.1 (reference default)

    def bbuild():
        torch.manual_seed(21)
        with torch.device(dev):
            mm = mbert.BertForSequenceClassification(bcfg)
        with torch.no_grad():
            for _, p in mm.named_parameters():
                if p.dim() >= 2:
                    p.normal_(0, 0.02)
        return mm.train()

    bg = torch.Generator().manual_seed(555 + rank)
    bids = torch.randint(1, 30522, (4, 512), generator=bg).to(dev)
    blab = torch.randint(0, 28, (4,), generator=bg).to(dev)
    bmask = torch.ones(4, 512, device=dev)
    bmask[1, 300:] = 0
    bseg = torch.zeros(4, 512, dtype=torch.long, device=dev)
    bpos = torch.arange(512, device=dev)

    def bloss(model):
        F.manual_dropout_seed(4000 + rank)  # every rank its own masks, the same ones for the local and the DDP run
        return torch.nn.functional.cross_entropy(model(bids, bmask, bseg, bpos).float(), blab)

    bb = bbuild(); bloss(bb).backward()
    bmean = {}
    for n, p in bb.named_parameters():
        t = p.grad.detach().clone(); dist.all_reduce(t); bmean[n] = t / world
    bm = bbuild()
    bddp = DistributedDataParallel(bm, device_ids=[local], comm="p2p", bucket_cap_mb=1)
    bloss(bddp).backward()
    torch.cuda.synchronize()
    worst = 0.0
    for n, p in bm.named_parameters():
        if n.endswith("k_linear.bias"):
            continue  # analytically zero (softmax shift invariance): rounding noise on both sides
        worst = max(worst, rel(p.grad, bmean[n]))
        check(same_on_all_ranks(p.grad), "BERT p2p grad %s differs across ranks" % n)
    check(worst <= 4e-3, "BERT (dropout 0.1, train) through DDP vs mean of local gradients: %g" % worst)
    res["bert_dropout_ddp_vs_mean"] = worst
    bddp.close()

    res["failures"] = FAILS
    allf = [None] * world
    dist.all_gather_object(allf, FAILS)
    res["failures_all_ranks"] = sorted({f for fl in allf for f in fl})
    if rank == 0:
        path = args.out or os.path.join(ROOT, "gpurun_out", "ddp_check_w%d.json" % world)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        json.dump(res, open(path, "w"), indent=1)
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()
    if res["failures_all_ranks"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
