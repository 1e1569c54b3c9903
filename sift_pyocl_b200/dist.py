"""Multi-GPU sharding of an image batch (no equivalent in the single-device reference).

One process per GPU (torchrun).  Image i of a batch goes to rank ``i % world``; each rank runs its own
SiftPlan; the ragged keypoint arrays are all-gathered over NCCL (``torch.distributed``): first the
int32 counts, then the records padded to the largest count.  The data path itself has no collective.
With the ``gloo`` backend (CPU tests) the same code moves the records through host tensors.
"""
import numpy
import torch
import torch.distributed as dist

from . import _lib


def shard_indices(n_images, rank, world):
    """Indices of the batch images owned by ``rank`` (round robin)."""
    return list(range(rank, n_images, world))


class _CudaView(object):
    """Expose a raw device pointer through __cuda_array_interface__ so torch can wrap it (no copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def device_records_tensor(plan, n):
    """uint8 torch tensor [n, 144] aliasing the plan's device-resident records of the last run."""
    ptr, _ = plan.device_records()
    if n == 0:
        return torch.empty((0, 144), dtype=torch.uint8, device="cuda:%d" % plan.device)
    t = torch.as_tensor(_CudaView(ptr, n * 144), device="cuda:%d" % plan.device)
    return t.view(n, 144)


class _PendingGather(object):
    """All-gather of ragged records split in two so that the host never waits for the exchange of the counts:
    begin (constructor) copies the local records and starts the all-gather of the counts asynchronously;
    ``finish()`` -- typically called one pipeline step later -- reads the counts and runs the all-gather of the
    records padded to the largest count."""

    def __init__(self, local, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.local = local.clone()  # the plan's record buffer is recycled by a later submit()
        dev = local.device
        self.cnt = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
        self.counts = torch.zeros(self.world, dtype=torch.int64, device=dev)
        self.work = dist.all_gather_into_tensor(self.counts, self.cnt, group=group, async_op=True)

    def finish(self):
        self.work.wait()
        counts_h = self.counts.cpu()
        nmax = max(int(counts_h.max()), 1)
        dev = self.local.device
        padded = torch.zeros((nmax, 144), dtype=torch.uint8, device=dev)
        padded[:self.local.shape[0]] = self.local
        gathered = torch.empty((self.world * nmax, 144), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(gathered, padded, group=self.group)
        gathered = gathered.view(self.world, nmax, 144)
        return [gathered[r, :int(counts_h[r])] for r in range(self.world)], counts_h


def allgather_records_begin(local, group=None):
    """Start the all-gather of ragged record tensors; returns a handle whose ``finish()`` completes it."""
    return _PendingGather(local, group)


def allgather_records(local, group=None):
    """All-gather ragged record tensors.

    :param local: uint8 tensor [n_local, 144] (CUDA for nccl, CPU for gloo)
    :return: (list of per-rank uint8 tensors [n_r, 144], int64 tensor of counts)
    """
    return _PendingGather(local, group).finish()


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
    t = torch.from_numpy(flat.view(numpy.uint8).reshape(-1, 144).copy()).to(dev)
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
