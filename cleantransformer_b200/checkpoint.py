"""Checkpoint I/O that does not stall the training step (SURVEY.md §8 f, N4).

The reference saves from inside the step loop, blocking: `torch.save(model.state_dict(), ...)` on rank 0 every
`save_interval` steps (examples/ft_bloom_DDP.py:155-156, ft_bloom.py:95-96) and, in the Trainer,
`_save_checkpoint` → `save_model` + `torch.save(optimizer.state_dict())` + scheduler / trainer state
(trainer/trainer.py:1303-1342, 1387-1404, 1412-1445). For Bloom-560M that is 2.2 GB of parameters (6.7 GB with
the AdamW moments) pulled over PCIe tensor by tensor and pickled while the GPU — and, under DDP, every other
rank waiting in its next all-reduce — is idle.

`AsyncCheckpointer.save()` returns as soon as the copies have been ENQUEUED:

  1. every tensor of the objects to save is found (nested dicts / lists / tuples) and grouped by the storage it
     lives in. The parameters and the moments of this package are views into flat arenas (`arena.py`), so a
     model + optimizer snapshot is three large copies, not 3 × 293 small ones; tensors that share memory
     (tied tables) stay shared in the file, exactly like `torch.save` of the live objects.
  2. stage "device" (default when HBM has room — 180 GB per GPU does): the ranges are copied device-to-device
     on the caller's stream (6.7 GB: about 2 ms of HBM traffic, not timed) into a staging buffer; the training stream may overwrite the
     live tensors right away. A side stream drains the staging buffer into pinned host memory.
     stage "host": the side stream reads the live tensors; `guard()` makes the caller's stream wait for that
     read and must be called before anything overwrites them (the next `optimizer.step()`; forward and
     backward do not write parameters or moments and overlap with the copy).
  3. a writer thread waits for the copy's event, then `torch.save`s to `<name>.tmp` and renames (a reader never
     sees a half-written file); host-only values (step counters, scheduler / trainer state, json) are deep-copied
     at call time. Pinned buffers are pooled and reused by the next snapshot.

The files are what the reference's loaders expect: plain `torch.save`d state_dicts (`pytorch_model.bin`,
`optimizer.pt`, `scheduler.pt`) and json. On CPU tensors (the host-logic tests) the same code runs without
streams: "copy" is a clone.
"""
import atexit
import copy
import json
import os
import queue
import threading
import weakref

import torch

_torch_save = torch.save        # the launcher's --ct-async-save replaces torch.save; the writer thread needs the real one
_LIVE = weakref.WeakSet()       # checkpointers that may hold a stage-"host" read of live tensors (guard_pending)


def _extent(t):
    """Number of storage elements from the tensor's first to one past its last addressed element."""
    if t.numel() == 0:
        return 0
    return 1 + sum((s - 1) * abs(st) for s, st in zip(t.shape, t.stride()))


class _Group:
    """All tensors of a snapshot that live in one (storage, dtype): one copy of the covering element range."""

    def __init__(self, dtype, device):
        self.dtype, self.device = dtype, device
        self.lo, self.hi = None, None
        self.members = []   # tensors

    def add(self, t):
        lo = t.storage_offset()
        hi = lo + _extent(t)
        self.lo = lo if self.lo is None else min(self.lo, lo)
        self.hi = hi if self.hi is None else max(self.hi, hi)
        self.members.append(t)

    def source(self):
        any_t = self.members[0]
        return torch.empty(0, dtype=self.dtype, device=self.device).set_(
            any_t.untyped_storage(), self.lo, (self.hi - self.lo,), (1,))


class _Placeholder:
    __slots__ = ("group", "shape", "stride", "offset")

    def __init__(self, group, t):
        self.group, self.shape, self.stride, self.offset = group, tuple(t.shape), tuple(t.stride()), t.storage_offset()


class AsyncCheckpointer:
    def __init__(self, stage="auto", pool_limit=2):
        if stage not in ("auto", "device", "host"):
            raise ValueError("stage must be 'auto', 'device' or 'host'")
        self.stage = stage
        self.pool_limit = pool_limit       # free pinned buffers kept per (dtype, numel)
        self._pool = {}
        self._pool_lock = threading.Lock()
        self._jobs = queue.Queue()
        self._error = None
        self._pending_read = None          # event of a stage-"host" copy still reading live tensors
        self._copy_stream = None
        self._thread = threading.Thread(target=self._writer, name="ct-checkpoint-writer", daemon=True)
        self._thread.start()
        self.saved = []                    # paths written so far (writer thread appends)
        _LIVE.add(self)

    # ---- snapshot ------------------------------------------------------------------------------
    def _collect(self, obj, groups):
        """Deep-copy `obj` with every tensor replaced by a placeholder; host-side leaves are copied now."""
        if torch.is_tensor(obj):
            t = obj.detach()
            if t.is_sparse or t.layout != torch.strided:
                raise TypeError("AsyncCheckpointer: only strided tensors can be saved")
            key = (t.untyped_storage().data_ptr() if t.numel() else id(t), t.dtype, str(t.device))
            g = groups.get(key)
            if g is None:
                g = groups[key] = _Group(t.dtype, t.device)
            g.add(t)
            return _Placeholder(g, t)
        if isinstance(obj, dict):
            new = copy.copy(obj)     # keeps the class and its attributes (a state_dict's `_metadata`)
            for k, v in obj.items():
                new[k] = self._collect(v, groups)
            return new
        if isinstance(obj, (list, tuple)):
            seq = [self._collect(v, groups) for v in obj]
            return seq if isinstance(obj, list) else tuple(seq)
        return copy.deepcopy(obj)

    @staticmethod
    def _materialise(obj, host_of):
        if isinstance(obj, _Placeholder):
            flat = host_of[id(obj.group)]
            return flat.as_strided(obj.shape, obj.stride, obj.offset - obj.group.lo)
        if isinstance(obj, dict):    # the skeleton is this snapshot's own copy: filled in place
            for k in obj:
                obj[k] = AsyncCheckpointer._materialise(obj[k], host_of)
            return obj
        if isinstance(obj, list):
            obj[:] = [AsyncCheckpointer._materialise(v, host_of) for v in obj]
            return obj
        if isinstance(obj, tuple):
            return tuple(AsyncCheckpointer._materialise(v, host_of) for v in obj)
        return obj

    def _host_buffer(self, dtype, numel, pinned):
        key = (dtype, numel, pinned)
        with self._pool_lock:
            free = self._pool.get(key)
            if free:
                return free.pop()
        return torch.empty(numel, dtype=dtype, device="cpu", pin_memory=pinned)   # explicit: run.py makes cuda the default device

    def _release(self, bufs):
        with self._pool_lock:
            for key, b in bufs:
                free = self._pool.setdefault(key, [])
                if len(free) < self.pool_limit:
                    free.append(b)

    def _stage_for(self, nbytes, device):
        if self.stage != "auto":
            return self.stage
        free, _ = torch.cuda.mem_get_info(device)
        return "device" if free > 2 * nbytes + (1 << 30) else "host"

    def save(self, files, json_files=None, on_done=None):
        """`files`: {path: object} written with torch.save (objects may hold tensors at any depth);
        `json_files`: {path: json-able object}; `on_done()`: called by the writer thread once all of them are on
        disk (completion markers, rotation of older checkpoints). Returns after the copies are enqueued."""
        self._raise_pending_error()
        groups = {}
        skeleton = {path: self._collect(obj, groups) for path, obj in files.items()}
        json_files = {p: copy.deepcopy(o) for p, o in (json_files or {}).items()}
        cuda_groups = [g for g in groups.values() if g.device.type == "cuda" and g.hi > g.lo]
        host_of, held, event = {}, [], None
        if cuda_groups:
            dev = cuda_groups[0].device
            nbytes = sum((g.hi - g.lo) * g.members[0].element_size() for g in cuda_groups)
            stage = self._stage_for(nbytes, dev)
            cur = torch.cuda.current_stream(dev)
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(dev)
            side = self._copy_stream
            srcs = []
            for g in cuda_groups:
                src = g.source()
                if stage == "device":
                    staged = torch.empty_like(src)      # on the caller's stream: reusable by it once freed
                    staged.copy_(src)
                    staged.record_stream(side)
                    src = staged
                srcs.append(src)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for g, src in zip(cuda_groups, srcs):
                    key = (g.dtype, g.hi - g.lo, True)
                    host = self._host_buffer(g.dtype, g.hi - g.lo, True)
                    host.copy_(src, non_blocking=True)
                    host_of[id(g)] = host
                    held.append((key, host))
                event = torch.cuda.Event()
                event.record(side)
            if stage == "host":
                self._pending_read = event
            del srcs
        for g in groups.values():
            if id(g) not in host_of:    # host tensors (step counters, CPU models of the tests) and empty ones: cloned now
                host_of[id(g)] = g.source().clone() if g.hi > g.lo else torch.empty(0, dtype=g.dtype, device="cpu")
        self._jobs.put((event, skeleton, json_files, host_of, held, on_done))

    def guard(self):
        """Make the current stream wait until a stage-"host" snapshot has finished READING the live tensors.
        Call before the next in-place update of what was saved (optimizer.step()); free in stage "device"."""
        ev, self._pending_read = self._pending_read, None
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    # ---- writer thread -------------------------------------------------------------------------
    @staticmethod
    def _atomic(path, write):
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        tmp = path + ".tmp"
        write(tmp)
        os.replace(tmp, path)

    def _writer(self):
        while True:
            job = self._jobs.get()
            if job is None:
                self._jobs.task_done()
                return
            event, skeleton, json_files, host_of, held, on_done = job
            try:
                if event is not None:
                    event.synchronize()     # blocks this thread only
                for path, obj in skeleton.items():
                    real = self._materialise(obj, host_of)
                    self._atomic(path, lambda tmp, real=real: _torch_save(real, tmp))
                    self.saved.append(path)
                for path, obj in json_files.items():
                    def w(tmp, obj=obj):
                        with open(tmp, "w") as f:
                            json.dump(obj, f, indent=2, sort_keys=True)
                            f.write("\n")
                    self._atomic(path, w)
                    self.saved.append(path)
                if on_done is not None:
                    on_done()
            except BaseException as e:  # surfaced by the next save() / wait()
                self._error = e
            finally:
                self._release(held)
                self._jobs.task_done()

    def _raise_pending_error(self):
        e, self._error = self._error, None
        if e is not None:
            raise RuntimeError("an earlier asynchronous checkpoint failed: %r" % (e,)) from e

    def wait(self):
        """Block until everything handed to save() is on disk; re-raises a writer failure."""
        self._jobs.join()
        self._raise_pending_error()

    def close(self):
        if self._thread.is_alive():
            self._jobs.join()
            self._jobs.put(None)
            self._thread.join()
        self._raise_pending_error()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def unwrap(model):
    """The module whose state_dict the reference saves (`model.module` under DistributedDataParallel,
    examples/ft_bloom_DDP.py:147-150)."""
    return model.module if hasattr(model, "module") and isinstance(model.module, torch.nn.Module) else model


def guard_pending():
    """Called by this package's optimizers at the top of step(): no parameter or moment is overwritten while a
    stage-"host" snapshot is still reading it, whoever drives the loop."""
    for ck in list(_LIVE):
        if ck._pending_read is not None:
            ck.guard()


_default = None


def default_checkpointer():
    """Process-wide checkpointer, flushed at interpreter exit."""
    global _default
    if _default is None:
        _default = AsyncCheckpointer()
        atexit.register(_default.close)
    return _default


def has_device_tensor(obj, _depth=0):
    if torch.is_tensor(obj):
        return obj.is_cuda
    if isinstance(obj, dict):
        return any(has_device_tensor(v, _depth + 1) for v in obj.values())
    if isinstance(obj, (list, tuple)):
        return any(has_device_tensor(v, _depth + 1) for v in obj)
    return False


def save_async(obj, path):
    """`torch.save(obj, path)` that returns once the copies are enqueued (the launcher's `--ct-async-save` routes the
    reference scripts' `torch.save(model.state_dict(), ...)`, examples/ft_bloom_DDP.py:155-156, through here)."""
    default_checkpointer().save({os.fspath(path): obj})
