"""Multi-GPU predictions: shard the draws, replicate the tables, gather the results.

Every draw is independent and the tables are read-only, so the path shards with no data-path
collective: rank r of W evaluates the contiguous slice
``[B r / W, B (r + 1) / W)`` of the draws on its own replica of the table, and the per-rank result
slabs ``[B_r, 1 + R]`` are collected with ONE ``gather`` (NCCL over NVLink on GPUs, gloo in the CPU
tests).  One process per GPU, launched with ``torch.distributed.run``.

When the caller wants the results in HOST memory (the reference's API returns numpy arrays), a
device-side gather funnels every rank's rows through rank ``dst``'s PCIe link.  On one node
``predict_batch_sharded(..., gather='host')`` avoids that: the ranks share one POSIX
shared-memory segment, registered with CUDA in every process, and each GPU copies its own rows
into its slice over its own PCIe link; the only collective left is a barrier.
"""

import atexit

import numpy as np

_SEGMENTS = {}   # (group id, n_bytes) -> SharedHostArray


def shard_bounds(n_draws, rank, world_size):
    """Contiguous slice of ``n_draws`` owned by ``rank``: sizes differ by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError('rank {} outside world of size {}'.format(rank, world_size))
    return n_draws * rank // world_size, n_draws * (rank + 1) // world_size


def shard_params(params, rank, world_size):
    """Slice a dict of ``[B]`` arrays (or a ``[B, k]`` array) to this rank's draws."""
    if isinstance(params, dict):
        n_draws = max(np.shape(v)[0] for v in params.values() if np.ndim(v) > 0)
        lo, hi = shard_bounds(n_draws, rank, world_size)
        return {k: (v[lo:hi] if np.ndim(v) > 0 else v) for k, v in params.items()}
    lo, hi = shard_bounds(len(params), rank, world_size)
    return params[lo:hi]


def gather_rows(local, n_total, dst=0, group=None):
    """Collect the per-rank row slabs of a ``[n_total, C]`` result on rank ``dst``.

    ``local`` is this rank's ``[hi - lo, C]`` tensor for ``shard_bounds(n_total, rank, world)``.
    Returns the concatenated ``[n_total, C]`` tensor on ``dst`` and ``None`` elsewhere.  Slabs are
    padded to a common row count because ``gather`` needs equal shapes.
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = max(shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0]
               for r in range(world))
    if local.shape[0] != rows:
        padded = torch.zeros((rows,) + tuple(local.shape[1:]), dtype=local.dtype,
                             device=local.device)
        padded[:local.shape[0]] = local
    else:
        padded = local.contiguous()
    slabs = ([torch.empty_like(padded) for _ in range(world)] if rank == dst else None)
    dist.gather(padded, slabs, dst=dst, group=group)
    if rank != dst:
        return None
    pieces = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        pieces.append(slabs[r][:hi - lo])
    return torch.cat(pieces, dim=0)


def gather_slab_chunks(compute_chunk, slab, full, n_chunks=3, dst=0, group=None):
    """The data-parallel step with its one collective overlapped: this rank's ``[n_local, C]``
    result slab is produced in ``n_chunks`` row ranges -- ``compute_chunk(c0, c1)`` launches the
    kernels that write ``slab[c0:c1]`` on the current stream (``TabCorr.predict_into_slab`` does,
    through the output strides of the C ABI: no packing copies) -- and every range is sent to
    rank ``dst`` with an asynchronous ``gather`` as soon as it is launched, so that the transfer
    of range c runs while the kernels of range c + 1 do.  On ``dst`` the receive buffers are the
    row blocks of ``full`` (``[world * n_local, C]``, rank-major like ``shard_bounds``), i.e. the
    result is assembled in place, without a concatenation.  All ranks hold equally many rows.
    Returns after every gather has completed (on the current stream for NCCL)."""
    import torch.distributed as dist
    n_local = slab.shape[0]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        compute_chunk(0, n_local)
        if full is not None and full.data_ptr() != slab.data_ptr():
            full.copy_(slab)
        return
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_chunks = max(1, min(int(n_chunks), n_local)) if n_local else 1
    cuts = [n_local * c // n_chunks for c in range(n_chunks + 1)]
    works = []
    for c in range(n_chunks):
        c0, c1 = cuts[c], cuts[c + 1]
        if c1 > c0:
            compute_chunk(c0, c1)
        receive = None
        if rank == dst:
            receive = [full[r * n_local + c0:r * n_local + c1] for r in range(world)]
        works.append(dist.gather(slab[c0:c1], receive, dst=dst, group=group, async_op=True))
    for work in works:
        work.wait()


class OverlappedGather:
    """Steps of a data-parallel job whose one collective runs behind the NEXT step's kernels.

    A step's result rows leave for rank ``dst`` with an asynchronous ``gather`` the moment its
    kernels are queued; the following step's kernels (written into the other of ``depth`` slabs)
    run while NVLink moves them, so in steady state a step costs ``max(kernels, gather)`` instead
    of their sum.  ``dst`` assembles every step in place in one of ``depth`` ``[world * n_local,
    width]`` buffers (rank-major rows like ``shard_bounds``).  Use::

        pipe = OverlappedGather(n_local, width, device)
        for step in steps:
            slab = pipe.begin()          # this step's [n_local, width] slab, free to overwrite
            halotab.predict_into_slab(theta, slab)
            full = pipe.submit()         # dst: the buffer the step lands in (valid after wait)
        pipe.finish()                    # every gather has completed on the current stream

    ``begin`` makes the current stream wait for the gather that last read the slab (``depth``
    steps ago); a consumer on ``dst`` calls ``wait(step)`` before reading that step's buffer.
    Without an initialised process group the slab is its own result."""

    def __init__(self, n_local, width, device, dst=0, group=None, depth=2, dtype=None):
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.group, self.dst, self.depth = group, dst, max(2, int(depth))
        self.active = dist.is_available() and dist.is_initialized() and \
            dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0
        self.n_local = int(n_local)
        dtype = dtype or torch.float64
        self.slabs = [torch.empty((n_local, width), dtype=dtype, device=device)
                      for _ in range(self.depth)]
        self.full = None
        if self.active and self.rank == dst:
            self.full = [torch.empty((self.world * n_local, width), dtype=dtype, device=device)
                         for _ in range(self.depth)]
        self.works = {}
        self.step = 0

    def begin(self):
        """The slab of the step about to be computed (its previous gather is waited for)."""
        self.wait(self.step - self.depth)
        return self.slabs[self.step % self.depth]

    def submit(self):
        """Queue the gather of the slab ``begin`` handed out; returns ``dst``'s landing buffer."""
        k = self.step
        self.step += 1
        slab = self.slabs[k % self.depth]
        if not self.active:
            return slab
        receive, full = None, None
        if self.rank == self.dst:
            full = self.full[k % self.depth]
            n = self.n_local
            receive = [full[r * n:(r + 1) * n] for r in range(self.world)]
        self.works[k] = self.dist.gather(slab, receive, dst=self.dst, group=self.group,
                                         async_op=True)
        return full

    def wait(self, step):
        work = self.works.pop(step, None)
        if work is not None:
            work.wait()

    def finish(self):
        for k in sorted(self.works):
            self.wait(k)


class PeerCopyGather:
    """:class:`OverlappedGather` with the copy engines as transport: every rank copies its step's
    rows into rank ``dst``'s result slab (``PeerSlab``: CUDA IPC mapping, NVLink / NVSwitch) with
    an asynchronous device-to-device ``copy_`` on a side stream.  An NCCL gather is a kernel; the
    fused prediction kernel is persistent and owns every SM's registers and shared memory, so a
    gather queued behind step k only runs in the gap before step k + 1 and delays it (measured on
    8 B200: 4.04 ms per step against 4.06 ms for the blocking gather).  A DMA copy needs no SM:
    it runs while the next step's kernels do.

    Same protocol: ``begin()`` -> slab to fill, ``submit()``, ``finish()``; ``finish`` also
    holds a barrier, after which ``landed(step)`` on ``dst`` is complete.  One node only."""

    def __init__(self, n_local, width, device, dst=0, group=None, depth=2):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group, self.dst, self.depth = group, dst, max(2, int(depth))
        self.active = dist.is_available() and dist.is_initialized() and \
            dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0
        self.device = torch.device(device)
        self.n_local = int(n_local)
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.slabs = [torch.empty((n_local, width), dtype=torch.float64, device=self.device)
                      for _ in range(self.depth)]
        self.peers, self.mine = [], []
        if self.active:
            for _ in range(self.depth):
                peer = PeerSlab(self.world * n_local, width, dst=dst, group=group, device=index)
                self.peers.append(peer)
                self.mine.append(peer.rows_tensor(self.rank * n_local, (self.rank + 1) * n_local))
        self.copy_stream = torch.cuda.Stream(self.device)
        self.done = {}
        self.step = 0

    def begin(self):
        event = self.done.pop(self.step - self.depth, None)
        if event is not None:
            self.torch.cuda.current_stream(self.device).wait_event(event)
        return self.slabs[self.step % self.depth]

    def submit(self):
        torch = self.torch
        k = self.step
        self.step += 1
        slab = self.slabs[k % self.depth]
        if not self.active:
            return slab
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        self.copy_stream.wait_event(ready)
        with torch.cuda.stream(self.copy_stream):
            self.mine[k % self.depth].copy_(slab, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        self.done[k] = done
        return self.landed(k)

    def landed(self, step):
        """``dst``: the ``[world * n_local, width]`` tensor step ``step`` lands in."""
        if not self.active:
            return self.slabs[step % self.depth]
        return self.peers[step % self.depth].tensor

    def finish(self):
        stream = self.torch.cuda.current_stream(self.device)
        for k in sorted(self.done):
            stream.wait_event(self.done.pop(k))
        if self.active:
            stream.synchronize()
            self.dist.barrier(group=self.group)

    def close(self):
        self.finish()
        self.mine = []
        for peer in self.peers:
            peer._row_holders = []
            peer.close()
        self.peers = []


class SlabRows:
    """Rows ``[lo, hi)`` of a :class:`PeerSlab` as seen from this process: a raw device pointer
    (possibly into another GPU's memory), the row count and the row width in doubles."""

    def __init__(self, ptr, n_rows, width):
        self.ptr, self.n_rows, self.width = int(ptr), int(n_rows), int(width)


class PeerSlab:
    """The ``[n_total, width]`` float64 result slab of a sharded batch in the HBM of rank
    ``dst``, mapped into every rank of the node (CUDA IPC; peer access over NVLink / NVSwitch).

    The path's only collective is the collection of the per-rank result rows on one rank
    (SURVEY.md section 8(e)).  With a peer slab there is no separate gather: every rank hands
    ``rows(lo, hi)`` to ``TabCorr.predict_into_slab`` and the epilogue kernel of the prediction
    (``finalize_kernel``, coalesced row stores) writes the rank's results straight into rank
    ``dst``'s memory while the other SMs still compute; what is left is one barrier.  ``tensor``
    (on ``dst``) is the assembled result."""

    def __init__(self, n_total, width, dst=0, group=None, device=None):
        import ctypes
        import torch
        import torch.distributed as dist
        from . import _lib
        self.lib = _lib.load()
        self.n_total, self.width, self.dst, self.group = int(n_total), int(width), dst, group
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        n_bytes = max(8, 8 * self.n_total * self.width)
        self.base = ctypes.c_void_p()
        self.owner = self.rank == dst
        handle = [None]
        error = None
        if self.owner:
            raw = (ctypes.c_ubyte * 64)()
            try:
                _lib.check(self.lib.tc_peer_alloc(self.device, n_bytes, ctypes.byref(self.base),
                                                  raw))
                handle[0] = bytes(raw)
            except _lib.TabCorrB200Error as err:   # the other ranks learn it from the None handle
                error = err
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast_object_list(handle, src=dist.get_global_rank(group, dst) if group else dst,
                                       group=group)
            if not self.owner and handle[0] is not None:
                raw = (ctypes.c_ubyte * 64).from_buffer_copy(handle[0])
                try:
                    _lib.check(self.lib.tc_peer_open(self.device, raw, ctypes.byref(self.base)))
                except _lib.TabCorrB200Error as err:
                    error = err
            # every rank takes the same decision (a rank that raised alone would leave the
            # others waiting in the next collective)
            ok = torch.tensor([0 if (error is not None or handle[0] is None) else 1],
                              dtype=torch.int32, device=torch.device('cuda', self.device))
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                if self.base.value:
                    (self.lib.tc_peer_free if self.owner else self.lib.tc_peer_close)(
                        self.device, self.base)
                self.base = None
                raise RuntimeError('peer-mapped result slab unavailable on this node (CUDA IPC / '
                                   'peer access): {}'.format(error or 'failed on another rank'))
        elif error is not None:
            raise error
        self.tensor = None
        if self.owner:
            class _Holder:   # torch tensor over the raw allocation (dst only)
                pass
            holder = _Holder()
            holder.__cuda_array_interface__ = {
                'shape': (self.n_total, self.width), 'typestr': '<f8',
                'data': (int(self.base.value), False), 'version': 2}
            self._holder = holder
            self.tensor = torch.as_tensor(holder, device=torch.device('cuda', self.device))

    def rows(self, lo, hi):
        return SlabRows(int(self.base.value) + 8 * int(lo) * self.width, hi - lo, self.width)

    def rows_tensor(self, lo, hi):
        """Rows ``[lo, hi)`` as a CUDA tensor of THIS process (on ranks other than ``dst`` the
        memory behind it is rank ``dst``'s, reached through the peer mapping): the destination of
        an asynchronous ``copy_``, which the copy engines move over NVLink without any SM."""
        import torch

        class _Holder:
            pass
        holder = _Holder()
        holder.__cuda_array_interface__ = {
            'shape': (int(hi - lo), self.width), 'typestr': '<f8',
            'data': (int(self.base.value) + 8 * int(lo) * self.width, False), 'version': 2}
        self._row_holders = getattr(self, '_row_holders', []) + [holder]
        return torch.as_tensor(holder, device=torch.device('cuda', self.device))

    def close(self):
        if self.base is None or not self.base.value:
            return
        import torch.distributed as dist
        self.tensor = None
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.barrier(group=self.group)   # nobody writes any more
        if self.owner:
            self.lib.tc_peer_free(self.device, self.base)
        else:
            self.lib.tc_peer_close(self.device, self.base)
        self.base = None


class SharedHostArray:
    """A float64 host buffer all ranks of one node map (``multiprocessing.shared_memory``), page
    locked for CUDA in every process (``cudaHostRegister``) so that device-to-host copies into it
    are asynchronous DMA transfers."""

    def __init__(self, n_doubles, group=None):
        import torch
        import torch.distributed as dist
        from multiprocessing import shared_memory
        rank = dist.get_rank(group)
        n_bytes = max(8, 8 * int(n_doubles))
        name = [None]
        if rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=n_bytes)
            name[0] = self.shm.name
        dist.broadcast_object_list(name, src=dist.get_global_rank(group, 0) if group else 0,
                                   group=group)
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
            try:   # the creator unlinks; keep Python's resource tracker from doing it twice
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, 'shared_memory')
            except Exception:
                pass
        self.owner = rank == 0
        self.array = np.ndarray((n_bytes // 8,), dtype=np.float64, buffer=self.shm.buf)
        self.tensor = torch.from_numpy(self.array)
        self.registered = False
        if torch.cuda.is_available():
            status = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), n_bytes, 0)
            self.registered = int(status) == 0
        dist.barrier(group=group)   # every rank has mapped the segment before anyone may unlink

    def close(self):
        import torch
        if self.shm is None:
            return
        if self.registered:
            try:
                torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            except Exception:
                pass
        self.tensor = self.array = None
        # unlink first and on its own: close() raises BufferError while a caller still holds views
        # of the segment (they stay valid: the mapping lives until the last view is gone), and
        # that must not leave the /dev/shm entry behind
        if self.owner:
            try:
                self.shm.unlink()
            except Exception:
                pass
        try:
            self.shm.close()
        except Exception:
            pass
        self.shm = None


_SEGMENT_CALLS = {}


def _shared_segment(n_doubles, group=None):
    """Two segments per group, used alternately: the views rank ``dst`` got from call k stay
    untouched during call k + 1 (whose writers are only ordered against dst by that call's final
    barrier) and are reused by call k + 2, when dst's caller has long started call k + 1 -- so
    "valid until the next call" holds without an extra barrier at entry."""
    calls = _SEGMENT_CALLS.get(id(group), 0)
    _SEGMENT_CALLS[id(group)] = calls + 1
    key = (id(group), int(n_doubles), calls & 1)
    if key not in _SEGMENTS:
        for old_key in [k for k in _SEGMENTS if k[0] == key[0] and (k[1] != key[1] or k[2] == key[2])]:
            _SEGMENTS.pop(old_key).close()      # segments of another batch size are dropped
        _SEGMENTS[key] = SharedHostArray(n_doubles, group)
    return _SEGMENTS[key]


@atexit.register
def _close_segments():
    for segment in list(_SEGMENTS.values()):
        segment.close()
    _SEGMENTS.clear()


def single_node(group=None):
    """True when all ranks of the group run on the same host."""
    import socket
    import torch.distributed as dist
    names = [None] * dist.get_world_size(group)
    dist.all_gather_object(names, socket.gethostname(), group=group)
    return len(set(names)) == 1


def shared_host_fits(n_doubles, group=None):
    """True when /dev/shm has room for a segment of ``n_doubles`` float64 (decided on rank 0 and
    broadcast, so that all ranks take the same path; touching pages of an over-committed tmpfs
    segment would raise SIGBUS) or when an equal segment is already mapped."""
    import os
    import torch.distributed as dist
    if any(k[:2] == (id(group), int(n_doubles)) for k in _SEGMENTS):
        return True
    fits = [True]
    if dist.get_rank(group) == 0:
        try:
            stat = os.statvfs('/dev/shm')
            fits[0] = stat.f_bavail * stat.f_frsize >= 8 * int(n_doubles) + (16 << 20)
        except OSError:
            fits[0] = False
    dist.broadcast_object_list(fits, src=dist.get_global_rank(group, 0) if group else 0,
                               group=group)
    return bool(fits[0])


def _predict_batch_shared_host(halotab, params, n_total, n_gauss_prim, model, dst, group,
                               predict_kwargs):
    """Every rank evaluates its slice host-to-host and writes the results into its rows of a
    shared, CUDA-registered host segment; ``dst`` returns views of it."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    shape = tuple(halotab.tpcf_shape)
    n_r = int(np.prod(shape))
    segment = _shared_segment(n_total * (1 + n_r), group)
    ngal_all = segment.tensor[:n_total].view(n_total, 1)
    xi_all = segment.tensor[n_total:n_total * (1 + n_r)].view(n_total, n_r, 1)
    lo, hi = shard_bounds(n_total, rank, world)
    if hi > lo:
        local = shard_params(params, rank, world)
        if hasattr(halotab, '_predict_batch_pipelined'):
            halotab.predict_batch(local, n_gauss_prim=n_gauss_prim, model=model,
                                  out=(ngal_all[lo:hi], xi_all[lo:hi]), **predict_kwargs)
        else:   # Interpolator, TableSet, test stand-ins: results to the host, then into the slice
            ngal, xi = halotab.predict_batch(local, n_gauss_prim=n_gauss_prim, model=model,
                                             as_numpy=False, **predict_kwargs)
            ngal_all[lo:hi].copy_(ngal.reshape(-1, 1), non_blocking=True)
            xi_all[lo:hi].copy_(xi.reshape(hi - lo, n_r, 1), non_blocking=True)
            if ngal.is_cuda:
                torch.cuda.current_stream(ngal.device).synchronize()
    dist.barrier(group=group)   # every slice has landed in host memory
    if rank != dst:
        return None
    return ngal_all.numpy()[:, 0], xi_all.numpy().reshape((n_total,) + shape)


def predict_batch_sharded(halotab, params, n_gauss_prim=10, model=None, dst=0, group=None,
                          n_chunks=3, gather='nccl', **predict_kwargs):
    """``halotab.predict_batch`` over all ranks of the process group.

    ``params`` holds ALL draws on every rank (dict of ``[B]`` arrays or ``[B, k]`` array); each
    rank evaluates its slice and rank ``dst`` receives ``(ngal [B], xi [B, *tpcf_shape])`` as numpy
    arrays (other ranks get ``None``).  Works for ``TabCorr`` and ``Interpolator`` instances.

    The slice is evaluated in ``n_chunks`` pieces: the gather of piece c (asynchronous, NCCL) and,
    on ``dst``, the device-to-host copy of the gathered piece c - 1 overlap the kernels of the
    next piece, so that rank ``dst``'s PCIe link -- which carries every rank's results -- is busy
    during the computation instead of after it.  Results do not depend on ``n_chunks``.

    ``gather``: 'nccl' (default) is that device-side gather; 'host' (one node only) lets every
    rank copy its rows into a shared, CUDA-registered host segment over its own PCIe link, with a
    barrier as the only collective -- the arrays ``dst`` gets back are views of that segment and
    stay valid until the next call (two segments alternate, so the next call's writers cannot
    touch them); 'auto' picks 'host' when all ranks share a node.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if isinstance(params, dict):
        n_total = max(np.shape(v)[0] for v in params.values() if np.ndim(v) > 0)
    else:
        n_total = len(params)
    if gather not in ('nccl', 'host', 'auto'):
        raise ValueError("gather must be 'nccl', 'host' or 'auto'")
    if world > 1 and gather != 'nccl' and (gather == 'host' or single_node(group)):
        n_doubles = n_total * (1 + int(np.prod(halotab.tpcf_shape)))
        if shared_host_fits(n_doubles, group):
            return _predict_batch_shared_host(halotab, params, n_total, n_gauss_prim, model, dst,
                                              group, predict_kwargs)
        # /dev/shm too small for the result: the device-side gather below needs no host segment
    local = shard_params(params, rank, world)
    lo_rank, hi_rank = shard_bounds(n_total, rank, world)
    n_local = hi_rank - lo_rank
    rows_max = max(shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0]
                   for r in range(world))
    n_chunks = max(1, min(int(n_chunks), rows_max)) if world > 1 else 1
    # piece c of every rank covers the same local row range [c0, c1) (clipped to the rank's rows)
    cuts = [rows_max * c // n_chunks for c in range(n_chunks + 1)]

    def piece(c):
        c0, c1 = min(cuts[c], n_local), min(cuts[c + 1], n_local)
        if isinstance(local, dict):
            return {k: (v[c0:c1] if np.ndim(v) > 0 else v) for k, v in local.items()}, c1 - c0
        return local[c0:c1], c1 - c0

    host, shape, pending = None, None, None
    use_cuda = torch.cuda.is_available()
    copy_stream = torch.cuda.Stream() if (use_cuda and rank == dst and world > 1) else None

    def drain(entry):
        """On dst: wait for gather c, then copy every rank's rows into the pinned host array."""
        work, slabs, c = entry
        if rank != dst or copy_stream is None:
            work.wait()
            if rank != dst:
                return
        else:
            # only the copy stream waits for the gather: the compute stream keeps launching
            with torch.cuda.stream(copy_stream):
                work.wait()
        for r in range(world):
            lo_r, hi_r = shard_bounds(n_total, r, world)
            c0, c1 = min(cuts[c], hi_r - lo_r), min(cuts[c + 1], hi_r - lo_r)
            if c1 <= c0:
                continue
            if copy_stream is not None:
                with torch.cuda.stream(copy_stream):
                    host[lo_r + c0:lo_r + c1].copy_(slabs[r][:c1 - c0], non_blocking=True)
                slabs[r].record_stream(copy_stream)
            else:
                host[lo_r + c0:lo_r + c1].copy_(slabs[r][:c1 - c0])

    for c in range(n_chunks):
        sub, n_sub = piece(c)
        if n_sub > 0:
            ngal, xi = halotab.predict_batch(sub, n_gauss_prim=n_gauss_prim, model=model,
                                             as_numpy=False, **predict_kwargs)
            shape = tuple(xi.shape[1:])
            slab = torch.cat([ngal.reshape(-1, 1), xi.reshape(xi.shape[0], -1)], dim=1)
        else:
            slab = None
        if world == 1:
            from .tabcorr import _to_host
            full = _to_host(slab)
            return full[:, 0], full[:, 1:].reshape((n_total,) + shape)
        if shape is None:   # a rank without rows still takes part in the gathers
            shape = tuple(halotab.tpcf_shape)
        width = 1 + int(np.prod(shape))
        rows = cuts[c + 1] - cuts[c]
        device = slab.device if slab is not None else (
            torch.device('cuda', torch.cuda.current_device()) if use_cuda else torch.device('cpu'))
        if slab is None or slab.shape[0] != rows:
            padded = torch.zeros((rows, width), dtype=torch.float64, device=device)
            if slab is not None:
                padded[:slab.shape[0]] = slab
            slab = padded
        if rank == dst and host is None:
            host = torch.empty((n_total, width), dtype=torch.float64, pin_memory=use_cuda)
        slabs = [torch.empty_like(slab) for _ in range(world)] if rank == dst else None
        work = dist.gather(slab.contiguous(), slabs, dst=dst, group=group, async_op=True)
        if pending is not None:
            drain(pending)
        pending = (work, slabs, c)
    drain(pending)
    if rank != dst:
        return None
    if copy_stream is not None:
        copy_stream.synchronize()
    full = host.numpy()
    return full[:, 0], full[:, 1:].reshape((n_total,) + shape)
