"""Multi-GPU sharding of an image batch (no equivalent in the single-device reference).

One process per GPU (torchrun).  Image i of a batch goes to rank ``i % world``; each rank runs its own
SiftPlan; the ragged keypoint arrays are all-gathered over NCCL (``torch.distributed``).  The data path
itself has no collective.  With the ``gloo`` backend (CPU tests) the same code moves the records through
host tensors.

Two forms of the exchange:
  * ``allgather_records`` -- blocking, exact sizes: counts first, then the records padded to the largest count;
  * ``RecordExchange`` -- the pipelined form used per step by ``bench.py`` and ``keypoints_batch``: a fixed-capacity
    slab per rank whose first row carries the count, gathered by ONE collective on a side stream, so the host
    never waits for the exchange of the counts and nothing is allocated per step.
"""
import numpy
import torch
import torch.distributed as dist

from . import _lib

REC = 144  # bytes per dtype_kp record


def shard_indices(n_images, rank, world):
    """Indices of the batch images owned by ``rank`` (round robin)."""
    return list(range(rank, n_images, world))


class _CudaView(object):
    """Expose a raw device pointer through __cuda_array_interface__ so torch can wrap it (no copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def device_records_tensor(plan, n):
    """uint8 torch tensor [n, 144] aliasing the plan's device-resident records of the last run.

    The memory belongs to the plan and is recycled by a later submit(): consume it on a torch stream and then
    call ``plan.wait_stream(that_stream)`` (RecordExchange does), or clone it and synchronise."""
    ptr, _ = plan.device_records()
    if n == 0:
        return torch.empty((0, REC), dtype=torch.uint8, device="cuda:%d" % plan.device)
    t = torch.as_tensor(_CudaView(ptr, n * REC), device="cuda:%d" % plan.device)
    return t.view(n, REC)


def allgather_records(local, group=None):
    """All-gather ragged record tensors (blocking).

    :param local: uint8 tensor [n_local, 144] (CUDA for nccl, CPU for gloo)
    :return: (list of per-rank uint8 tensors [n_r, 144], int64 tensor of counts)
    """
    world = dist.get_world_size(group)
    dev = local.device
    cnt = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, cnt, group=group)
    counts_h = counts.cpu()
    nmax = max(int(counts_h.max()), 1)
    padded = torch.zeros((nmax, REC), dtype=torch.uint8, device=dev)
    padded[:local.shape[0]] = local
    gathered = torch.empty((world * nmax, REC), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    gathered = gathered.view(world, nmax, REC)
    return [gathered[r, :int(counts_h[r])] for r in range(world)], counts_h


class _Pending(object):
    """One exchange in flight; owns references to the slabs it was started with (the exchange may enlarge its
    buffers while an older step is still pending)."""

    def __init__(self, ex, send, recv, capacity, work, event, n_local, spill, counts_host=None):
        self.ex, self.send, self.recv, self.capacity = ex, send, recv, capacity
        self.work, self.event, self.n_local, self.spill = work, event, n_local, spill
        self.counts_host = counts_host

    def finish(self):
        """(list of per-rank uint8 tensors [n_r, 144], int64 counts).  The tensors alias the exchange's receive
        buffer: valid until the second begin() from now (two buffers alternate)."""
        ex = self.ex
        if self.work is not None:
            self.work.wait()
        if self.event is not None:
            self.event.synchronize()
        recv = self.recv
        if self.counts_host is not None:
            counts = self.counts_host.view(torch.int64).reshape(-1).clone()
        else:
            counts = recv[:, 0, :8].contiguous().view(torch.int64).reshape(-1).cpu()
        if int(counts.max()) > self.capacity:
            # a rank produced more records than a slab holds (every rank sees the same counts, so all take this
            # branch together): exact-size blocking exchange of the full local arrays, and larger slabs from now on
            local = self.spill if self.spill is not None else self.send[1:1 + self.n_local]
            out = allgather_records(local, ex.group)
            if ex.capacity < int(counts.max()):
                ex._allocate(int(int(counts.max()) * 1.25) + 1024)
            return out
        return [recv[r, 1:1 + int(counts[r])] for r in range(ex.world)], counts


class RecordExchange(object):
    """Per-step all-gather of ragged keypoint records without a host round trip for the counts.

    Every rank contributes a slab of ``capacity + 1`` rows of 144 bytes: row 0 carries the count (int64), rows
    1.. the records.  ``begin()`` copies the local records into the send slab and starts ONE all-gather on a side
    stream; ``finish()`` -- typically one pipeline step later -- reads the gathered counts, which the copy engine has
    placed in page-locked host memory right after the collective (no kernel, no host wait beyond the event).  All
    buffers are allocated once (``capacity`` must be the same on every rank); a step whose count exceeds the
    capacity falls back to the exact-size exchange and enlarges the slabs.
    """

    def __init__(self, capacity, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.stream = torch.cuda.Stream(self.device) if self.cuda else None
        self.turn = 0
        self._allocate(int(capacity))

    def _allocate(self, capacity):
        self.capacity = capacity
        rows = capacity + 1
        self.send = [torch.zeros((rows, REC), dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.recv = [torch.empty((self.world, rows, REC), dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.header = [torch.zeros(1, dtype=torch.int64, pin_memory=self.cuda) for _ in range(2)]
        self.counts_host = [torch.zeros((self.world, 8), dtype=torch.uint8, pin_memory=self.cuda) for _ in range(2)]

    def begin(self, local, plan=None):
        """Start the exchange of ``local`` (uint8 [n, 144]; may alias a plan's record buffer: pass ``plan`` so that
        the plan does not recycle that buffer before the copy out of it has run)."""
        b = self.turn
        self.turn ^= 1
        n = int(local.shape[0])
        send, recv = self.send[b], self.recv[b]
        self.header[b][0] = n  # page-locked: the copy below is asynchronous; rewritten two steps later at the earliest
        header = self.header[b].view(torch.uint8)
        spill = None
        if not self.cuda:
            send[0, :8] = header
            send[1:1 + min(n, self.capacity)] = local[:self.capacity]
            if n > self.capacity:
                spill = local.clone()
            work = dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=self.group, async_op=True)
            return _Pending(self, send, recv, self.capacity, work, None, n, spill)
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            send[0, :8].copy_(header, non_blocking=True)
            send[1:1 + min(n, self.capacity)].copy_(local[:self.capacity], non_blocking=True)
            if n > self.capacity:
                spill = local.clone()
            if plan is not None:  # the plan's next writer of this record buffer waits for the copies above
                if hasattr(plan, "hold_records"):
                    plan.hold_records(self.stream.cuda_stream)
                else:
                    plan.wait_stream(self.stream.cuda_stream)
            dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=self.group)
            # the gathered counts travel to page-locked host memory by the copy engine: reading them in finish()
            # then needs no kernel (a kernel, however small, queues behind the plan's persistent CTAs for SM space
            # and stalled the host thread by up to a kernel's duration)
            counts_host = self.counts_host[b]
            for r in range(self.world):
                counts_host[r].copy_(recv[r, 0, :8], non_blocking=True)
            event = torch.cuda.Event()
            event.record(self.stream)
        return _Pending(self, send, recv, self.capacity, None, event, n, spill, counts_host)


def allgather_records_begin(local, group=None, exchange=None, plan=None):
    """Start the all-gather of ragged record tensors; returns a handle whose ``finish()`` completes it.
    Without an ``exchange`` a one-shot RecordExchange sized for ``local`` is used (allocates)."""
    if exchange is None:
        world = dist.get_world_size(group)
        cap = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=group)
        exchange = RecordExchange(max(int(cap.item()), 1), local.device, group)
        del world
    return exchange.begin(local, plan)


def records_to_numpy(t):
    """uint8 tensor [n, 144] -> numpy.recarray of dtype_kp."""
    a = t.contiguous().cpu().numpy()
    return a.reshape(-1).view(_lib.dtype_kp).view(numpy.recarray)


def keypoints_batch(plan, images, group=None, gather=True):
    """Keypoints of a batch sharded over the ranks of ``group``.

    :param plan: this rank's SiftPlan
    :param images: the FULL batch (sequence of arrays, or a callable ``i -> image``); only the images
                   owned by this rank are touched
    :return: list over the batch of recarrays (every rank gets every image's keypoints when
             ``gather`` is true, else only its own entries are filled)
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_images = len(images)
    mine = shard_indices(n_images, rank, world)
    get = images if callable(images) else images.__getitem__
    # the rank's images go through the plan's pipelined form when it has one (copies of one image overlap the
    # kernels of the others), else one at a time
    many = getattr(plan, "keypoints_many", None)
    local = list(many(get(i) for i in mine)) if many is not None else [plan.keypoints(get(i)) for i in mine]
    sizes = [kp.size for kp in local]
    out = [None] * n_images
    if not gather:
        for i, kp in zip(mine, local):
            out[i] = kp
        return out
    use_cuda = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda:%d" % plan.device) if use_cuda else torch.device("cpu")
    flat = numpy.concatenate(local) if local else numpy.zeros(0, _lib.dtype_kp)
    t = torch.from_numpy(flat.view(numpy.uint8).reshape(-1, REC).copy()).to(dev)
    per_rank, _ = allgather_records(t, group)
    # per-image sizes of every rank
    rounds = (n_images + world - 1) // world
    sz = torch.zeros(rounds, dtype=torch.int64, device=dev)
    sz[:len(sizes)] = torch.tensor(sizes, dtype=torch.int64)
    all_sz = torch.zeros(world * rounds, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_sz, sz, group=group)
    all_sz = all_sz.view(world, rounds).cpu().numpy()
    for r in range(world):
        recs = records_to_numpy(per_rank[r])
        start = 0
        for j, i in enumerate(shard_indices(n_images, r, world)):
            out[i] = recs[start:start + all_sz[r, j]]
            start += all_sz[r, j]
    return out


def match_sharded(match_plan, kp1, kp2, group=None):
    """MatchPlan.match(kp1, kp2, raw_results=True) with the rows of ``kp1`` sharded over the ranks of ``group``
    (SURVEY.md 8f rank 4): rank r matches kp1[r::world] against all of kp2, the index pairs are all-gathered.

    Every rank passes the same kp1 / kp2 and receives the same int32 [m, 2] array, sorted by the kp1 index (the
    reference's output order is nondeterministic: atomic_inc, matching_gpu.cl:100).
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = numpy.arange(rank, kp1.shape[0], world)
    local = match_plan.match(kp1[mine], kp2, raw_results=True) if mine.size else numpy.zeros((0, 2), numpy.int32)
    local = numpy.ascontiguousarray(local, numpy.int32).copy()
    if local.shape[0]:
        local[:, 0] = mine[local[:, 0]]  # back to indices into the full kp1
    use_cuda = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda:%d" % match_plan.device) if use_cuda else torch.device("cpu")
    t = torch.from_numpy(local).to(dev)
    cnt = torch.tensor([t.shape[0]], dtype=torch.int64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, cnt, group=group)
    counts_h = counts.cpu()
    nmax = max(int(counts_h.max()), 1)
    padded = torch.zeros((nmax, 2), dtype=torch.int32, device=dev)
    padded[:t.shape[0]] = t
    gathered = torch.empty((world * nmax, 2), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    gathered = gathered.view(world, nmax, 2).cpu().numpy()
    out = numpy.concatenate([gathered[r, :int(counts_h[r])] for r in range(world)])
    return out[numpy.argsort(out[:, 0], kind="stable")]
