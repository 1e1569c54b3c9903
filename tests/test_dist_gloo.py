"""N>1 host logic on CPU: world_size-2 gloo run of the batch sharding + ragged all-gather
(sift_pyocl_b200/dist.py).  The plan is replaced by a stub so no GPU is needed; on the GPU box the same
code runs over NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _StubPlan(object):
    device = 0

    def keypoints(self, image):
        from sift_pyocl_b200._lib import dtype_kp
        seed, n = int(image[0]), int(image[1])
        rng = np.random.default_rng(seed)
        kp = np.zeros(n, dtype_kp)
        kp["x"], kp["y"] = rng.random(n), rng.random(n)
        kp["desc"] = rng.integers(0, 256, (n, 128))
        return kp.view(np.recarray)


class _StubPipelinedPlan(_StubPlan):
    """A plan with the pipelined generator form (SiftPlan.keypoints_many)."""
    calls = 0

    def keypoints_many(self, images):
        for im in images:
            type(self).calls += 1
            yield self.keypoints(im)


def _worker(rank, world, port, n_images, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sift_pyocl_b200 import dist as sdist
    images = [np.array([100 + i, (7 * i) % 5 * 3]) for i in range(n_images)]  # some images have 0 keypoints
    out = sdist.keypoints_batch(_StubPlan(), images)
    ok = len(out) == n_images
    for i, kp in enumerate(out):
        want = _StubPlan().keypoints(images[i])
        ok = ok and kp.size == want.size and np.array_equal(kp.x, want.x) and np.array_equal(kp.desc, want.desc)
    mine = sdist.shard_indices(n_images, rank, world)
    piped = sdist.keypoints_batch(_StubPipelinedPlan(), images)  # same result through keypoints_many
    ok = ok and _StubPipelinedPlan.calls == len(mine)
    ok = ok and all(a.size == b.size and np.array_equal(a.desc, b.desc) for a, b in zip(out, piped))
    local_only = sdist.keypoints_batch(_StubPlan(), images, gather=False)
    ok = ok and all((local_only[i] is not None) == (i in mine) for i in range(n_images))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [5, 2, 1])
def test_sharded_batch_allgather_world2(n_images):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


def _pipelined_worker(rank, world, port, q):
    """Two-phase all-gather as the pipelined e2e loop uses it: begin(step i+1) before finish(step i)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sift_pyocl_b200 import dist as sdist
    steps = [torch.full((3 * i + rank, 144), 10 * i + rank, dtype=torch.uint8) for i in range(4)]
    pending, ok = None, True
    results = []
    for t in steps:
        started = sdist.allgather_records_begin(t)
        t.fill_(255)  # the source buffer is recycled right away, like a plan slot
        if pending is not None:
            results.append(pending.finish())
        pending = started
    results.append(pending.finish())
    for i, (per_rank, counts) in enumerate(results):
        for r in range(world):
            ok = ok and int(counts[r]) == 3 * i + r and per_rank[r].shape == (3 * i + r, 144)
            ok = ok and bool((per_rank[r] == 10 * i + r).all())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _exchange_worker(rank, world, port, q):
    """RecordExchange: fixed-capacity slabs, one collective per step, overflow falls back to the exact exchange."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sift_pyocl_b200 import dist as sdist
    ex = sdist.RecordExchange(8, "cpu")
    sizes = [3, 0, 8, 20, 5, 31, 2]          # steps 3 and 5 exceed the capacity in force
    steps = [torch.full((sz + rank, 144), 7 * i + rank, dtype=torch.uint8) for i, sz in enumerate(sizes)]
    pending, results, ok = None, [], True

    def done(p):  # the returned tensors alias the receive slabs (valid until the second begin() from now): copy
        per_rank, counts = p.finish()
        results.append(([t.clone() for t in per_rank], counts))
    for t in steps:
        started = ex.begin(t)
        t.fill_(255)
        if pending is not None:
            done(pending)
        pending = started
    done(pending)
    for i, (per_rank, counts) in enumerate(results):
        for r in range(world):
            ok = ok and int(counts[r]) == sizes[i] + r and tuple(per_rank[r].shape) == (sizes[i] + r, 144)
            ok = ok and bool((per_rank[r] == 7 * i + r).all())
    ok = ok and ex.capacity > 8              # enlarged after the overflow
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_record_exchange_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


def test_pipelined_allgather_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pipelined_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


class _StubMatch(object):
    device = 0

    def match(self, kp1, kp2, raw_results=True):  # nearest x, accepted when the x values agree
        out = [(i, int(np.argmin(abs(kp2.x - a)))) for i, a in enumerate(kp1.x) if abs(kp2.x - a).min() < 1e-6]
        return np.array(out, np.int32).reshape(-1, 2)


def _match_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sift_pyocl_b200 import dist as sdist
    from sift_pyocl_b200._lib import dtype_kp
    rng = np.random.default_rng(5)
    k1 = np.zeros(37, dtype_kp).view(np.recarray)
    k2 = np.zeros(29, dtype_kp).view(np.recarray)
    k1.x = rng.permutation(37)
    k2.x = rng.permutation(45)[:29]
    got = sdist.match_sharded(_StubMatch(), k1, k2)
    want = _StubMatch().match(k1, k2)
    q.put((rank, bool(np.array_equal(got, want[np.argsort(want[:, 0], kind="stable")]) and len(want) > 5)))
    dist.destroy_process_group()


def test_match_sharded_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_match_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


def test_shard_indices():
    from sift_pyocl_b200.dist import shard_indices
    assert shard_indices(64, 3, 8) == list(range(3, 64, 8))  # BASELINE config 3: 64 images over 8 GPUs
    assert sorted(sum((shard_indices(10, r, 4) for r in range(4)), [])) == list(range(10))
