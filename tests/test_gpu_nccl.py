"""2-rank NCCL checks on real GPUs (skipped on a single-GPU box): the records every rank receives from
dist.keypoints_batch / RecordExchange / the C-ABI siftb_allgather_kp equal the single-GPU records."""
import ctypes
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _images(n=5, size=384):
    from sift_pyocl_b200.utils import multiscale_image
    return [multiscale_image(size, 900 + i) for i in range(n)]


def _worker(rank, world, port, q):
    try:
        import torch
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        import sift_pyocl_b200 as sift
        from sift_pyocl_b200 import _lib, dist as sdist
        from helpers import same_records
        imgs = _images()
        plan = sift.SiftPlan(shape=imgs[0].shape, dtype=np.float32, device=rank)
        solo = [plan.keypoints(im) for im in imgs]                 # single-GPU records of every image
        ok = True
        failed = []

        def check(name, cond):
            if not cond:
                failed.append(name)
            return bool(cond)
        # (a) the sharded batch with the NCCL gather: every rank ends up with every image's records
        out = sdist.keypoints_batch(plan, imgs)
        ok = check("keypoints_batch", len(out) == len(imgs) and all(same_records(a, b) for a, b in zip(out, solo))) and ok
        # (b) the pipelined per-step exchange on device-resident records (what bench.py does at N > 1), with a
        # capacity small enough that one step takes the overflow path
        mine = sdist.shard_indices(len(imgs), rank, world)
        theirs = sdist.shard_indices(len(imgs), 1 - rank, world)
        cap = sorted(k.size for k in solo)[len(solo) // 2]
        ex = sdist.RecordExchange(cap, "cuda:%d" % rank)
        pending, got = None, []
        for step in range(len(theirs) if len(theirs) < len(mine) else len(mine)):
            plan.submit(imgs[mine[step]])
            n = plan.collect(records=False)
            started = ex.begin(sdist.device_records_tensor(plan, n), plan)
            if pending is not None:
                got.append(pending.finish())
            pending = started
            # the plan's buffers are recycled immediately by further images
            plan.keypoints(imgs[mine[step]])
            plan.keypoints(imgs[(mine[step] + 1) % len(imgs)])
        got.append(pending.finish())
        for step, (per_rank, counts) in enumerate(got):
            for r in range(world):
                idx = sdist.shard_indices(len(imgs), r, world)[step]
                ok = check("exchange count step %d rank %d" % (step, r), int(counts[r]) == solo[idx].size) and ok
                ok = check("exchange records step %d rank %d" % (step, r), same_records(sdist.records_to_numpy(per_rank[r]), solo[idx])) and ok
        # (c) the C-ABI communicator: unique id from rank 0, all-gather of this rank's first image
        lib = _lib.load()
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            _lib.check(lib.siftb_comm_unique_id(buf))
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        comm = ctypes.c_void_p()
        _lib.check(lib.siftb_comm_init(rank, world, ctypes.c_char_p(bytes(uid.cpu().numpy().tobytes())), rank,
                                       ctypes.byref(comm)))
        plan.submit(imgs[mine[0]])
        n = plan.collect(records=False)
        recs, _ = plan.device_records()
        counts = np.zeros(world, np.int32)
        total = ctypes.c_int()
        cap_out = sum(k.size for k in solo)
        host = np.zeros(cap_out, _lib.dtype_kp)
        _lib.check(lib.siftb_allgather_kp(comm, ctypes.c_void_p(recs), n, counts.ctypes.data_as(_lib.c_int_p),
                                          _lib.ptr(host), cap_out, ctypes.byref(total)))
        first = [sdist.shard_indices(len(imgs), r, world)[0] for r in range(world)]
        ok = check("c-abi counts", counts.tolist() == [solo[i].size for i in first] and total.value == counts.sum()) and ok
        start = 0
        for r in range(world):
            ok = check("c-abi records rank %d" % r, same_records(host[start:start + counts[r]].view(np.recarray), solo[first[r]])) and ok
            start += counts[r]
        _lib.check(lib.siftb_comm_destroy(comm))
        # (d) MatchPlan with the rows of list 1 sharded over the ranks, index pairs all-gathered over NCCL
        mp = sift.MatchPlan(device=rank)
        other = np.concatenate([solo[1], solo[0][::-1]]).view(np.recarray)   # holds a copy of every keypoint of list 1
        whole = mp.match(solo[0], other, raw_results=True)
        whole = whole[np.argsort(whole[:, 0], kind="stable")]
        sharded = sdist.match_sharded(mp, solo[0], other)
        ok = check("match_sharded (%d vs %d pairs)" % (len(sharded), len(whole)), np.array_equal(sharded, whole) and len(whole) > 0) and ok
        q.put((rank, bool(ok), "; ".join(failed)))
        dist.destroy_process_group()
    except Exception as exc:  # report instead of hanging the parent
        import traceback
        q.put((rank, False, traceback.format_exc() + str(exc)))


def test_two_rank_nccl_gather_equals_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60)
    assert [(r, ok) for r, ok, _ in res] == [(0, True), (1, True)], res
