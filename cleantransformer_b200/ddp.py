"""DistributedDataParallel — the wrapper the reference's DDP example constructs
(`DDP(model, device_ids=[local_rank])`, examples/ft_bloom_DDP.py:99), built the way README.md:46-52
describes the hand-rolled version: (1) parameter sync at construction, (2) gradient buckets,
(3) bucket reduction overlapped with the rest of backward.

Call-compatible surface (what the example uses): ctor `(module, device_ids=[r])`, `.module`,
`__call__/forward(**kw)`, `parameters()`, `state_dict()` with the `module.` prefix
(ft_bloom_DDP.py:99,124,150,156), `train()/eval()`.

B200 design
  * all parameters are flattened into one arena (arena.py); its gradient buffer is the symmetric,
    peer-mapped buffer of csrc/comm.cu, so every wgrad kernel writes straight into NVLink-visible
    memory and buckets are plain [lo, hi) ranges of that buffer (no copies in or out);
  * buckets are laid out in reverse parameter order (~ the order gradients appear in backward),
    ~25 MiB each like torch's default; a bucket is launched on a side stream the moment all of its
    gradients have been enqueued (functional.grad_written) — the all-reduce kernel itself waits for
    the peers with flags in the same symmetric memory;
  * averaging (1/world) is fused into the all-reduce kernel;
  * where the driver offers it the symmetric buffer is a cuMemCreate allocation bound to an NVLink MULTICAST object
    (file descriptors passed between the ranks over a Unix socket): the bucket all-reduce then adds inside the
    NVSwitch (`multimem.ld_reduce` / `multimem.st`, csrc/comm.cu) and a rank only touches 1/W of a bucket — 16 small
    CTAs instead of 48 register-staged ones beside the persistent GEMMs of backward. Fallback: cudaMalloc + IPC
    handles, two-shot over unicast peer pointers;
  * collective epochs live in device memory, so a whole training step — buckets included — can be captured in a CUDA
    graph and replayed (graphs.GraphedTrainStep);
  * `comm="nccl"` keeps torch.distributed.all_reduce (NCCL on GPUs, gloo in the CPU tests) on the
    same bucket layout as the baseline/oracle.
Parameters used more than once per step (tied embedding / lm_head) are reduced when the last of
their gradients has been written (`_ct_expected_writes`, set by the tie helpers), in practice at the
end of backward — except on the P2P path when the second write is an embedding scatter
(`_ct_sparse_second_write`): the dense LM-head wgrad is all-reduced the moment it is enqueued (first
kernel of backward, fully hidden), and the token-row gradients are exchanged sparsely at the end
(`ct_embedding_bwd_allranks`: (W-1)*T*H*4 bytes per rank instead of an exposed 1 GB all-reduce).
"""
import ctypes
import os
import socket

import torch
import torch.distributed as dist

from . import _lib
from .arena import ParamArena

BUCKET_CAP_MB = 25


class _CudaView:
    """Non-owning __cuda_array_interface__ view of library-owned device memory."""

    def __init__(self, ptr, numel):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (ptr, False),
                                         "version": 2, "strides": None}


def plan_buckets(sizes_offsets, cap_elems, solo=()):
    """sizes_offsets: [(offset, numel_padded)] in arena order. Returns [(lo, hi, [param indices])]
    covering the arena back to front in chunks of about cap_elems. Indices in `solo` get a bucket of
    their own (parameters that are reduced on their own schedule)."""
    buckets, cur, cur_lo, cur_hi = [], [], None, None
    for idx in range(len(sizes_offsets) - 1, -1, -1):
        off, n = sizes_offsets[idx]
        if cur and (((cur_hi - off) > cap_elems and (cur_hi - cur_lo) > 0) or idx in solo or cur[-1] in solo):
            buckets.append((cur_lo, cur_hi, cur))
            cur, cur_lo, cur_hi = [], None, None
        if not cur:
            cur_hi = off + n
        cur.append(idx)
        cur_lo = off
    if cur:
        buckets.append((cur_lo, cur_hi, cur))
    return buckets


class DistributedDataParallel(torch.nn.Module):
    def __init__(self, module, device_ids=None, output_device=None, bucket_cap_mb=BUCKET_CAP_MB,
                 comm=None, process_group=None, max_ctas=48, **unused):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("DistributedDataParallel needs torch.distributed.init_process_group first "
                               "(it is only the bootstrap channel for the P2P path)")
        self.module = module
        self.group = process_group
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("DistributedDataParallel: module has no trainable parameters")
        self.device = params[0].device
        on_gpu = self.device.type == "cuda"
        self.comm = comm or os.environ.get("CT_DDP_COMM") or ("p2p" if on_gpu else "nccl")
        if self.comm not in ("p2p", "nccl"):
            raise ValueError("comm must be 'p2p' (peer-memory kernels) or 'nccl' (library collective), got %r" % self.comm)
        if self.comm == "p2p" and not on_gpu:
            raise RuntimeError("comm='p2p' needs CUDA parameters")
        self._peer_mem = self.comm == "p2p"
        self.nvls = False
        self.max_ctas = int(os.environ.get("CT_DDP_CTAS", 0)) or None  # None: 16 / 32 with NVLS (see _launch), else `max_ctas`
        self._max_ctas_unicast = max_ctas
        self.final_ctas = int(os.environ.get("CT_DDP_FINAL_CTAS", 148))
        grad_buf = None
        n_total = sum((p.numel() + 63) // 64 * 64 for p in params)
        # tied tables whose second gradient is an embedding scatter: dense half reduced early, sparse half
        # exchanged at the end of backward (P2P path only)
        self._sparse = [p for p in params if getattr(p, "_ct_expected_writes", 1) == 2 and
                        getattr(p, "_ct_sparse_second_write", False) and p.dim() == 2 and p.shape[1] % 4 == 0]
        if not (self._peer_mem and self.world > 1) or os.environ.get("CT_DDP_SPARSE_TIED", "1") == "0":
            self._sparse = []
        self._stage_cap = int(os.environ.get("CT_DDP_STAGE_TOKENS", 16384))
        if self._peer_mem and self.world > 1:
            n_stage = 0
            if self._sparse:
                hmax = max(p.shape[1] for p in self._sparse)
                n_stage = 64 + self._stage_cap * hmax + 2 * self._stage_cap  # header | rows f32 | ids int64
            full = self._init_p2p(n_total + n_stage)
            grad_buf = full[:n_total]
            if self._sparse:
                self._stage_hdr_off = n_total
                self._stage_rows_off = n_total + 64
                self._stage_ids_off = n_total + 64 + self._stage_cap * hmax
                self._stage_hdr = full[n_total:n_total + 2].view(torch.int64)
                self._stage_rows = full[self._stage_rows_off:self._stage_ids_off]
                self._stage_ids = full[self._stage_ids_off:self._stage_ids_off + 2 * self._stage_cap].view(torch.int64)
        self.arena = ParamArena(params, grad_buffer=grad_buf)
        # (1) parameter / buffer sync from rank 0 (README.md:47; torch DDP's _sync_module_states). With peer memory the flat
        # parameter arena travels through the (still unused) symmetric gradient buffer: ct_broadcast, no library collective
        if self._peer_mem and self.world > 1 and grad_buf is not None and self.arena.grad.numel() >= self.arena.flat.numel():
            n = self.arena.flat.numel()
            self.arena.grad[:n].copy_(self.arena.flat)
            st = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.load().ct_comm_barrier(st), "ct_comm_barrier")      # every rank has staged (rank 0's counts)
            _lib.check(_lib.load().ct_broadcast(0, n, 0, st), "ct_broadcast")
            self.arena.flat.copy_(self.arena.grad[:n])
            self.arena.grad.zero_()
            torch.cuda.synchronize(self.device)
        else:
            dist.broadcast(self.arena.flat, 0, group=self.group)
        for b in module.buffers():
            if b.numel():
                dist.broadcast(b, 0, group=self.group)
        for p in self.arena.params:
            p._ct_shadow_ver = -1
        # (2) buckets
        so = []
        for i, p in enumerate(self.arena.params):
            end = self.arena.offsets[i + 1] if i + 1 < len(self.arena.params) else self.arena.numel
            so.append((self.arena.offsets[i], end - self.arena.offsets[i]))
        sparse_ids = {id(p) for p in self._sparse}
        solo = {i for i, p in enumerate(self.arena.params) if id(p) in sparse_ids}
        self.buckets = plan_buckets(so, int(bucket_cap_mb * 1024 * 1024 // 4), solo)
        self._grad_off = {id(p): self.arena.offsets[i] for i, p in enumerate(self.arena.params) if id(p) in sparse_ids}
        self._early = {}
        self._scatter_how = {}
        self._bucket_of = {}
        for bi, (_, _, idxs) in enumerate(self.buckets):
            for i in idxs:
                self._bucket_of[id(self.arena.params[i])] = bi
        self._pending = None
        self._launched = None
        self._writes = None
        self._cb_queued = False
        self._comm_stream = torch.cuda.Stream(device=self.device, priority=-1) if on_gpu else None
        self.require_backward_grad_sync = True
        # (3) hooks: our kernels announce gradients through functional.grad_written; gradients that
        # come from torch autograd (plain nn modules) through post-accumulate hooks
        for p in self.arena.params:
            hooks = getattr(p, "_ct_grad_hooks", None)
            if hooks is None:
                p._ct_grad_hooks = hooks = []
            hooks.append(self._on_grad_written)
            p.register_post_accumulate_grad_hook(self._on_autograd_grad)
        for p in self._sparse:
            p._ct_embedding_bwd_override = self._sparse_embedding_bwd

    # ---------------------------------------------------------------- P2P bootstrap
    def _all_ok(self, ok):
        """Do ALL ranks agree that a bootstrap stage succeeded? (torch.distributed is the bootstrap channel.)"""
        t = torch.tensor([1 if ok else 0], device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return bool(int(t))

    def _exchange_fds(self, tag, send_fd, senders, receivers):
        """Pass open file descriptors between the local ranks over abstract Unix sockets (SCM_RIGHTS): every rank in
        `senders` sends `send_fd` to every OTHER rank in `receivers`. Returns {sender rank: fd} on the receivers."""
        base = "\0ct_b200_%s_%s_%s_" % (os.environ.get("MASTER_PORT", "0"), os.getuid(), tag)
        got = {}
        srv = None
        if self.rank in receivers:
            srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            srv.bind(base + str(self.rank))
            srv.listen(self.world)
        dist.barrier(group=self.group)  # every listener is up
        try:
            if self.rank in senders:
                for q in receivers:
                    if q == self.rank:
                        continue
                    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                    c.connect(base + str(q))
                    socket.send_fds(c, [bytes([self.rank])], [send_fd])
                    c.close()
            if srv is not None:
                for _ in [q for q in senders if q != self.rank]:
                    conn, _addr = srv.accept()
                    msg, fds, _flags, _a = socket.recv_fds(conn, 16, 1)
                    got[msg[0]] = fds[0]
                    conn.close()
        finally:
            if srv is not None:
                srv.close()
        dist.barrier(group=self.group)
        return got

    def _init_p2p(self, n_total):
        lib = _lib.load()
        nbytes = n_total * 4
        local = ctypes.c_void_p()
        everyone = list(range(self.world))
        vmm_ok, mc_ok = ctypes.c_int(0), ctypes.c_int(0)
        if os.environ.get("CT_DDP_VMM", "1") != "0":
            lib.ct_comm_vmm_supported(self.device.index, ctypes.byref(vmm_ok), ctypes.byref(mc_ok))
        if self._all_ok(bool(vmm_ok.value)):
            fd = ctypes.c_int(-1)
            ok = lib.ct_comm_vmm_init(self.rank, self.world, self.device.index, nbytes, ctypes.byref(local),
                                      ctypes.byref(fd)) == 0
            err = "" if ok else _lib.last_error()
            if self._all_ok(ok):
                peer = self._exchange_fds("mem", fd.value, everyone, everyone)
                arr = (ctypes.c_int * self.world)(*[peer.get(q, -1) for q in range(self.world)])
                ok = lib.ct_comm_vmm_connect(arr) == 0
                err = "" if ok else _lib.last_error()
                for f in list(peer.values()) + [fd.value]:
                    os.close(f)
                if self._all_ok(ok):
                    self._init_multicast(lib, bool(mc_ok.value))
                    return torch.as_tensor(_CudaView(local.value, n_total), device=self.device)
            if self.rank == 0 and err:
                print("[cleantransformer_b200.ddp] VMM peer memory unavailable (%s): falling back to IPC handles" % err,
                      flush=True)
            lib.ct_comm_finalize()
        dh = ctypes.create_string_buffer(64)
        sh = ctypes.create_string_buffer(64)
        _lib.check(lib.ct_comm_init(self.rank, self.world, self.device.index, nbytes, ctypes.byref(local), dh, sh),
                   "ct_comm_init")
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (dh.raw, sh.raw), group=self.group)
        dhs = b"".join(g[0] for g in gathered)
        shs = b"".join(g[1] for g in gathered)
        _lib.check(lib.ct_comm_connect(dhs, shs), "ct_comm_connect")
        dist.barrier(group=self.group)
        return torch.as_tensor(_CudaView(local.value, n_total), device=self.device)

    def _init_multicast(self, lib, supported):
        """Bind every rank's allocation to one NVLink multicast object (NVLS). Any failure leaves the unicast VMM
        mappings in place and the all-reduce on its two-shot unicast kernel."""
        want = supported and os.environ.get("CT_DDP_NVLS", "1") != "0"
        if not self._all_ok(want):
            return
        mcfd = ctypes.c_int(-1)
        ok = True
        if self.rank == 0:
            ok = lib.ct_comm_mc_create(ctypes.byref(mcfd)) == 0
        if not self._all_ok(ok):
            if self.rank == 0:
                print("[cleantransformer_b200.ddp] no multicast object (%s): two-shot unicast all-reduce" % _lib.last_error(),
                      flush=True)
            return
        got = self._exchange_fds("mc", mcfd.value, [0], list(range(self.world)))
        if self.rank != 0:
            ok = lib.ct_comm_mc_import(got[0]) == 0
            os.close(got[0])
        else:
            os.close(mcfd.value)
        if not self._all_ok(ok):
            return
        ok = lib.ct_comm_mc_add_device() == 0
        if not self._all_ok(ok):  # (also the barrier cuMulticastBindMem needs: every device has been added)
            if self.rank == 0:
                print("[cleantransformer_b200.ddp] cuMulticastAddDevice failed (%s)" % _lib.last_error(), flush=True)
            return
        ok = lib.ct_comm_mc_bind() == 0
        err = "" if ok else _lib.last_error()
        if not self._all_ok(ok):
            if err:
                print("[cleantransformer_b200.ddp] rank %d: multicast bind failed (%s)" % (self.rank, err), flush=True)
            return  # a partially bound team is unusable: NO rank addresses it (self.nvls stays False everywhere)
        self.nvls = True

    # ---------------------------------------------------------------- forward
    def forward(self, *args, **kwargs):
        if torch.is_grad_enabled() and self.require_backward_grad_sync:
            self._begin_step()
        return self.module(*args, **kwargs)

    def _begin_step(self):
        if self._pending is None:  # the previous synchronised backward has finished (or this is the first step)
            for p in self.arena.params:
                p._ct_uses = 0     # functional.note_use counts this forward's uses = the writes to expect
        self._pending = [len(idxs) for (_, _, idxs) in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._writes = {}
        self._cb_queued = False
        self._early = {}
        self._scatter_how = {}

    # ---------------------------------------------------------------- gradient notifications
    def _on_autograd_grad(self, p):
        view = p._ct_grad_view
        if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
            view.copy_(p.grad)  # gradient produced by torch autograd outside the arena: move it in
            p.grad = view
        self._on_grad_written(p)

    def _on_grad_written(self, p):
        if self._pending is None or not self.require_backward_grad_sync:
            return
        if not self._cb_queued:
            self._cb_queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._finish_backward)
        n = self._writes.get(id(p), 0) + 1
        self._writes[id(p)] = n
        bi = self._bucket_of[id(p)]
        if id(p) in self._grad_off:
            # tied table on the P2P path. A dense first write (the LM-head wgrad, first kernel of backward) is
            # all-reduced at once; every later write must be a token scatter that _sparse_embedding_bwd exchanged
            # across ranks itself (any number of them: GPT looks tokens_embed up twice with segment_ids,
            # modeling_gpt.py:186-188). A scatter that arrives BEFORE any dense write stays local and the table takes
            # the generic route below (dense all-reduce once every expected write is in).
            how = self._scatter_how.pop(id(p), None)  # set by _sparse_embedding_bwd right before this notification
            state = self._early.get(id(p))
            if how == "exchanged":
                return
            if how is None and state is None and n == 1:
                self._early[id(p)] = "dense"
                self._pending[bi] -= 1
                self._launch(bi)
                return
            if state == "dense":
                raise RuntimeError("DistributedDataParallel: a dense gradient of a tied embedding table arrived "
                                   "after its early all-reduce was launched (set CT_DDP_SPARSE_TIED=0 to reduce the "
                                   "table once, densely, at the end of backward)")
            self._early[id(p)] = "local"
        if self._launched[bi]:
            raise RuntimeError("DistributedDataParallel: a gradient was written after its bucket had been "
                               "all-reduced (parameter used more often in backward than its forward announced)")
        expected = getattr(p, "_ct_uses", 0) or getattr(p, "_ct_expected_writes", 1)
        if n != expected:
            return
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi, final=False):
        if os.environ.get("CT_DDP_DEBUG"):
            self._dbg = getattr(self, "_dbg", [])
            self._dbg.append((bi, final, sum(self._writes.values())))
        if self._launched[bi] or self.world == 1:
            self._launched[bi] = True
            return
        self._launched[bi] = True
        lo, hi, _ = self.buckets[bi]
        if os.environ.get("CT_DDP_SKIP_COMM"):  # timing experiments only: gradients stay local (WRONG results)
            return
        if self.comm == "p2p":
            cur = torch.cuda.current_stream(self.device)
            self._comm_stream.wait_stream(cur)
            with torch.cuda.stream(self._comm_stream):
                # buckets launched after backward has finished have nothing to overlap with: use every SM
                if self.nvls:
                    # per rank 1/W of a bucket crosses the SM: 16 CTAs at 8 ranks (8 / 16 / 32: 40.4 / 40.2 / 40.7 ms per
                    # step, r02i), 32 below (2 ranks: 42.1 ms with 16, 40.6 with 32, r02e)
                    mode, ctas = 2, (64 if final else (self.max_ctas or (16 if self.world >= 8 else 32)))
                else:
                    mode, ctas = 3, (self.final_ctas if final else (self.max_ctas or self._max_ctas_unicast))
                _lib.check(_lib.load().ct_allreduce_bucket(lo, hi - lo, 1.0 / self.world, mode, ctas,
                                                           self._comm_stream.cuda_stream), "ct_allreduce_bucket")
        else:  # baseline / oracle path: library collective on the same bucket layout
            seg = self.arena.grad[lo:hi]
            if self._comm_stream is not None:
                self._comm_stream.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(self._comm_stream):
                    dist.all_reduce(seg, group=self.group)
                    seg.mul_(1.0 / self.world)
            else:
                dist.all_reduce(seg, group=self.group)
                seg.mul_(1.0 / self.world)

    def _sparse_embedding_bwd(self, param, ids, dout, grad, padding_idx):
        """EmbeddingFn's scatter for a tied table (functional.EmbeddingFn.backward). `grad` already holds the
        rank-averaged dense half when the early all-reduce has been launched for this step."""
        from . import ops
        if self._pending is None or not self.require_backward_grad_sync or self._early.get(id(param)) != "dense":
            # not inside a synchronised backward (no_sync / plain use), or the table saw no dense gradient
            # first: local scatter; with a pending dense reduction of this bucket it is picked up there
            ops.embedding_bwd(ids, dout, grad, padding_idx)
            if self._pending is not None:
                self._scatter_how[id(param)] = "local"
            return
        T, H = ids.numel(), dout.shape[-1]
        if grad.data_ptr() != param._ct_grad_view.data_ptr():
            raise RuntimeError("DistributedDataParallel: the tied table's .grad is not its arena view")
        V = param.shape[0]
        cur = torch.cuda.current_stream(self.device)
        self._comm_stream.wait_stream(cur)
        rows, flat_ids = dout.reshape(-1, H), ids.reshape(-1)
        with torch.cuda.stream(self._comm_stream):
            # more tokens than the staging area holds (CT_DDP_STAGE_TOKENS, sized at construction inside the symmetric
            # buffer): several exchanges. Like the buckets, this assumes every rank runs the same batch shape.
            for c0 in range(0, T, self._stage_cap):
                n = min(self._stage_cap, T - c0)
                self._stage_hdr.fill_(n)
                self._stage_rows[:n * H].copy_(rows[c0:c0 + n].reshape(-1))
                self._stage_ids[:n].copy_(flat_ids[c0:c0 + n])
                _lib.check(_lib.load().ct_embedding_bwd_allranks(
                    self._stage_hdr_off, self._stage_rows_off, self._stage_ids_off, self._grad_off[id(param)], H, V,
                    int(padding_idx), 1.0 / self.world, self.final_ctas * 2, self._comm_stream.cuda_stream),
                    "ct_embedding_bwd_allranks")
        if not torch.cuda.is_current_stream_capturing():  # (graph memory is static; the side stream joins before the end)
            dout.record_stream(self._comm_stream)
            ids.record_stream(self._comm_stream)
        self._scatter_how[id(param)] = "exchanged"

    def _finish_backward(self):
        # gradients that never arrived (unused parameters) or multi-use parameters still pending:
        # reduce whatever is left, in bucket order (identical on every rank)
        for bi in range(len(self.buckets)):
            if not self._launched[bi]:
                self._launch(bi, final=True)
        if self._comm_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)
        if os.environ.get("CT_DDP_DEBUG") and self.rank == 0:
            print("DDP launch order (bucket, final?, grads written so far):", getattr(self, "_dbg", [])[:60],
                  "pending", self._pending, flush=True)
            self._dbg = []
        self._pending = None

    # ---------------------------------------------------------------- misc surface
    def close(self):
        """Release the peer-mapped buffers (one communication context per process: call this before wrapping another
        model with comm='p2p'). The module's gradients move back to ordinary memory; parameters are untouched."""
        if self._peer_mem and self.world > 1 and self.arena is not None:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            for p in self.arena.params:
                if p.grad is not None and p.grad.data_ptr() == p._ct_grad_view.data_ptr():
                    p.grad = None
                p._ct_grad_view = None
                p._ct_arena = None
                hooks = getattr(p, "_ct_grad_hooks", None)
                if hooks and self._on_grad_written in hooks:
                    hooks.remove(self._on_grad_written)
                if getattr(p, "_ct_embedding_bwd_override", None) is not None:
                    p._ct_embedding_bwd_override = None
            self.arena = None
            _lib.check(_lib.load().ct_comm_finalize(), "ct_comm_finalize")
            dist.barrier(group=self.group)
            self._peer_mem = False

    def no_sync(self):
        ddp = self

        class _NoSync:
            def __enter__(self_inner):
                ddp.require_backward_grad_sync = False

            def __exit__(self_inner, *a):
                ddp.require_backward_grad_sync = True

        return _NoSync()

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            if name == "module":
                raise
            return getattr(self.module, name)
