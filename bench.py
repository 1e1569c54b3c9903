#!/usr/bin/env python
"""bench.py -- keypoints/sec of SiftPlan.keypoints() on 4096x4096 float32 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # the CPU arm (oracle port, all host threads)

A "step" is one pass of the hot path (plan.keypoints) over one synthetic image per GPU.  Workload =
BASELINE.json configs[1]: SiftPlan 4096x4096 float32, 3 octaves x 3 scales (par.OctaveMax = 3), images
= seeded "multiscale noise" (sift_pyocl_b200.utils.multiscale_image, seed 1234 + index).
  value : whole-job keypoints/s with the images already resident in HBM (device pointer input), records
          left on the device, CUDA events on the plan's stream around EXACTLY K software-pipelined steps,
          max over ranks.  The K-step region is measured --repeats times (default 5); the line reports the
          median region and lists all of them (`timed_regions`).
  e2e   : the same metric through the public API with HOST buffers: pinned host image -> H2D copy ->
          kernels -> D2H copy of the records -> numpy recarray, every step (SiftPlan.keypoints_many, three
          images in flight); `one_at_a_time` is the reference's own call, SiftPlan.keypoints(host image).
  roofline : the Gaussian pyramid convolution = ALL blur launches of a step (normalise + first blur, 8*W*H
          bytes; five blur+DoG launches per octave, 12*W_o*H_o bytes each).  achieved =
          algorithmic bytes of the average launch / its average duration (CUDA events on the plan's stream
          in a K-step region with profiling on) against MEASURED_PEAKS.json hbm_gbs; the octave-0 launches
          and the first blur are broken out; `traffic` = DRAM bytes of the average launch from the committed
          ncu --set full capture (profiles/).
  cpu_baseline : the oracle (CPU port of the reference kernels, OpenMP) on the same image, all host
          threads and one thread.
For N > 1 (torchrun) every rank runs the same per-GPU work on different images (weak scaling).  The images are
independent, so the data path has no collective and `value` / `e2e` have none; `with_gather` reports the same
regions with every step's keypoint records all-gathered to every rank over NCCL (dist.RecordExchange: one
collective per step on a side stream, completed one pipeline step later) -- SURVEY 8e asks for both.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SIZE = 4096
OCTAVES = 3
N_IMAGES = 2  # distinct images cycled per rank (working set per image ~1.2 GB >> 126 MB L2)
# dram__bytes_read.sum + dram__bytes_write.sum of all 16 blur launches of one step, summed from the committed
# ncu --set full capture of the final kernels of the round (see TRAFFIC_NOTE); bytes per step
TRAFFIC_PER_STEP = 1028.72e6
TRAFFIC_NOTE = ("sum over the 16 blur launches of one step in profiles/r02_ncu_full_step_4096_3oct.csv (ncu --set full of "
                "the round's final kernels), divided by 16; below the algorithmic 1455 MB because the planes of "
                "octaves 1 and 2 are still dirty in the 126 MB L2 when their launches end")


def _peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled through NVML every few ms during the timed regions
    (nvidia-smi is the fallback; it only manages a few samples per second)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.max_mhz = index, [], False, None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                if n is not None:
                    mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                    self.samples.append((mhz, bool(r & n.nvmlClocksEventReasonHwSlowdown),
                                         bool(r & n.nvmlClocksEventReasonHwThermalSlowdown),
                                         bool(r & n.nvmlClocksEventReasonSwThermalSlowdown),
                                         bool(r & n.nvmlClocksEventReasonSwPowerCap)))
                    time.sleep(0.004)
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                         timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(",")]
                    if len(f) >= 6 and f[0].isdigit():
                        self.max_mhz = int(f[1]) if f[1].isdigit() else self.max_mhz
                        self.samples.append((int(f[0]),) + tuple(x.lower().startswith("active") for x in f[2:6]))
                    time.sleep(0.05)
            except Exception:
                time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        sm = sorted(s[0] for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(s[1 + i] for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def _workload_name():
    return "SiftPlan %dx%d float32, %d octaves x 3 scales" % (SIZE, SIZE, OCTAVES)


def _config(world):
    """The workload description: IDENTICAL in both arms (the driver compares the dicts)."""
    return {"workload": _workload_name(), "images_per_step_per_gpu": 1,
            "image": "multiscale noise (sift_pyocl_b200.utils.multiscale_image), seeds 1234+",
            "l2": "inputs larger than L2: %d distinct 67 MB images cycled per GPU, ~1.2 GB of planes rewritten per step"
                  % N_IMAGES,
            "gather": "none on the data path (independent images, one per GPU and step); the all-gather of the "
                      "records is timed separately: with_gather" if world > 1 else "none (1 GPU)"}


def _cpu_leg(siftref, img, threads, reps):
    """keypoints/s of the oracle port on ``threads`` host threads (bounded sample: ``reps`` images)."""
    siftref.set_num_threads(threads)
    n, t0 = 0, time.perf_counter()
    for _ in range(reps):
        n += siftref.keypoints(img, octave_max=OCTAVES).size
    dt = time.perf_counter() - t0
    return n / dt, 1e3 * dt / reps


def run_reference(args):
    """CPU arm: the oracle port of the reference's kernels on all host threads (the reference's own
    OpenCL path cannot run: no PyOpenCL / ICD in the image)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import siftref
    from sift_pyocl_b200.utils import multiscale_image
    siftref.set_num_threads(os.cpu_count())  # torchrun exports OMP_NUM_THREADS=1
    cores = siftref.num_threads()
    img = multiscale_image(SIZE, seed=1234)
    for _ in range(args.warmup):
        siftref.keypoints(img, octave_max=OCTAVES)
    nkp, t0 = 0, time.perf_counter()
    for _ in range(args.steps):
        nkp += siftref.keypoints(img, octave_max=OCTAVES).size
    dt = time.perf_counter() - t0
    v = nkp / dt
    line = {"impl": "reference", "metric": "keypoints/sec on 4096x4096 float32", "value": v, "unit": "keypoints/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(int(os.environ.get("WORLD_SIZE", "1"))),
            "cpu_baseline": {"value": v, "unit": "keypoints/s", "cores": cores, "kind": "port",
                             "sample": "%d x one %dx%d image (seed 1234), OpenMP on %d threads" % (args.steps, SIZE, SIZE, cores)},
            "e2e": {"value": v, "unit": "keypoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def blur_bytes_per_step(plan):
    """Algorithmic bytes of the blur family in one step and of the dominant (octave 0) launches."""
    total = 8 * plan.shape[0] * plan.shape[1]  # first blur: read image, write G0
    for (w, h) in plan.scales:
        total += 5 * 12 * int(w) * int(h)  # five blur+DoG launches per octave
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--repeats", type=int, default=5, help="the K-step timed region is measured this many times; "
                    "the line reports the median region (every region times exactly K steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    # stdout carries exactly one JSON line: libraries that chat on fd 1 (NCCL prints its version banner there at
    # any NCCL_DEBUG level >= VERSION) are sent to stderr, the line itself goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import sift_pyocl_b200 as sift
    from sift_pyocl_b200 import _lib, dist as sdist
    from sift_pyocl_b200.utils import multiscale_image

    sift.par["OctaveMax"] = OCTAVES
    plan = sift.SiftPlan(shape=(SIZE, SIZE), dtype=np.float32, device=local_rank)
    sift.par["OctaveMax"] = 100000
    stream = torch.cuda.ExternalStream(plan.queue, device=local_rank)

    # synthetic inputs: host (pinned) and device copies
    host_imgs, dev_imgs = [], []
    for i in range(N_IMAGES):
        img = multiscale_image(SIZE, seed=1234 + rank * N_IMAGES + i)
        pinned = _lib.pinned_empty(img.shape, np.float32)
        pinned[...] = img
        host_imgs.append(pinned)
        dev_imgs.append(torch.from_numpy(img).cuda())
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up; at N > 1 the ranks also agree on the slab size of the per-step record exchange
    # (pipelined like the timed steps: the plan allocates the planes of its second compute lane the first time two
    # images are in flight together)
    n_max = 0
    plan.submit(dev_imgs[0])
    for i in range(args.warmup):
        if i + 1 < args.warmup:
            plan.submit(dev_imgs[(i + 1) % N_IMAGES])
        n_max = max(n_max, plan.collect(records=False))
    exchange = None
    if world > 1:
        cap = torch.tensor([n_max], dtype=torch.int64, device="cuda")
        dist.all_reduce(cap, op=dist.ReduceOp.MAX)
        exchange = sdist.RecordExchange(int(int(cap.item()) * 1.12) + 1024, "cuda:%d" % local_rank)
        for i in range(2):  # NCCL warm-up of the exchange itself
            plan.submit(dev_imgs[i % N_IMAGES])
            n = plan.collect(records=False)
            exchange.begin(sdist.device_records_tensor(plan, n), plan).finish()

    host_x = [0.0]

    def device_region(profile, gather=False):
        """K steps, device-resident input, records left on the device; software-pipelined: three images are in
        flight (the plan's two compute lanes + one queued), the exchange of step i (N > 1) is completed one step
        later."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nkp, blur_ms, blur0_ms, first_ms, stage_ms = 0, 0.0, 0.0, 0.0, {}
        barrier()
        ev0.record(stream)
        t0 = time.perf_counter()
        pending = None
        # images in flight (3 = what SiftPlan.keypoints_many keeps; with the all-gather's host work in the loop, 2 left
        # a lane idle now and then: 1.89 instead of 1.82 ms per step at 2 GPUs)
        depth = int(os.environ.get("SIFTB_BENCH_DEPTH", "3"))
        for j in range(min(depth, args.steps)):
            plan.submit(dev_imgs[j % N_IMAGES])
        for i in range(args.steps):
            n = plan.collect(records=False)
            nkp += n
            started = None
            if exchange is not None and gather:
                th = time.perf_counter()
                started = exchange.begin(sdist.device_records_tensor(plan, n), plan)
                host_x[0] += time.perf_counter() - th
            events = plan.fetch_events() if profile else ()
            # (the next submit recycles the slot of the image just collected: its records and events are taken first)
            if i + depth < args.steps:
                plan.submit(dev_imgs[(i + depth) % N_IMAGES])
            if started is not None:
                th = time.perf_counter()
                if pending is not None:
                    pending.finish()
                pending = started
                host_x[0] += time.perf_counter() - th
            if profile:
                for name, ms in events:
                    key = name.split(" octave")[0]
                    stage_ms[key] = stage_ms.get(key, 0.0) + ms
                    if "blur" in name:
                        blur_ms += ms
                    if name == "blur + DoG octave 0":
                        blur0_ms += ms
                    if name == "normalize + init blur":
                        first_ms += ms
        cur = torch.cuda.current_stream()
        cur.wait_stream(stream)
        if exchange is not None and gather:
            cur.wait_stream(exchange.stream)  # the last exchange belongs to the timed region
        ev1.record(cur)
        if pending is not None:
            pending.finish()
        barrier()
        wall = time.perf_counter() - t0
        return ev0.elapsed_time(ev1) / 1e3, nkp, wall, blur_ms, blur0_ms, first_ms, stage_ms

    def e2e_region(gather=False):
        """K steps through the public API with HOST buffers: pinned image -> H2D -> kernels -> D2H records ->
        numpy recarray every step (SiftPlan.keypoints_many keeps three images in flight)."""
        barrier()
        e2e_kp, d2h, t0 = 0, 0, time.perf_counter()
        pending = None
        for kp in plan.keypoints_many(host_imgs[i % N_IMAGES] for i in range(args.steps)):
            e2e_kp += kp.size
            d2h += kp.size * 144 + 4 * (1 + 13 * plan.octave_max + 4)
            if exchange is not None and gather:
                started = exchange.begin(sdist.device_records_tensor(plan, kp.size), plan)
                if pending is not None:
                    pending.finish()
                pending = started
        if pending is not None:
            pending.finish()
        barrier()
        return time.perf_counter() - t0, e2e_kp, d2h

    def sync_region():
        """the same, strictly one image at a time: SiftPlan.keypoints(host image), the reference's call"""
        barrier()
        n, t0 = 0, time.perf_counter()
        for i in range(args.steps):
            n += plan.keypoints(host_imgs[i % N_IMAGES]).size
        barrier()
        return time.perf_counter() - t0, n

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = plan.launches
    regions = [device_region(False) for _ in range(max(args.repeats, 1))]
    launches = (plan.launches - launches0) // max(args.repeats, 1)
    # throughput without and with the gather.  The images are independent, so the data path has no
    # collective (`value`); the all-gather of every step's records to every rank is an extra, timed separately
    withgather = [device_region(False, gather=True) for _ in range(3)] if world > 1 else None
    plan.set_profile(True)   # stage breakdown + roofline launches: a separate region (event pairs cost ~1 %)
    prof = device_region(True)
    plan.set_profile(False)
    for kp in plan.keypoints_many(host_imgs[i % N_IMAGES] for i in range(3)):
        pass
    e2e_regions = [e2e_region() for _ in range(max(min(args.repeats, 3), 1))]
    e2e_gather = [e2e_region(gather=True) for _ in range(2)] if world > 1 else None
    sync_s, sync_kp = sync_region()
    sampler.stop_flag = True

    def median_by(rs, key):
        rs = sorted(rs, key=key)
        return rs[len(rs) // 2]
    dev_s, nkp, wall = median_by(regions, lambda r: r[0])[:3]
    e2e_s, e2e_kp, d2h = median_by(e2e_regions, lambda r: r[0])
    all_dev_s = [r[0] for r in regions]
    wg_s = median_by(withgather, lambda r: r[0])[0] if withgather else 0.0
    wg_e2e_s = median_by(e2e_gather, lambda r: r[0])[0] if e2e_gather else 0.0
    if world > 1:
        t = torch.tensor([dev_s, e2e_s, sync_s, float(nkp), float(e2e_kp), float(launches), float(sync_kp), wg_s, wg_e2e_s],
                         dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dev_s, e2e_s, sync_s, wg_s, wg_e2e_s = (float(tmax[0]), float(tmax[1]), float(tmax[2]), float(tmax[7]),
                                                float(tmax[8]))
        nkp, e2e_kp, launches, sync_kp = float(t[3]), float(t[4]), int(t[5]), float(t[6])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = _peaks()
    _, _, _, blur_ms, blur0_ms, first_ms, stage_ms = prof
    bbytes = blur_bytes_per_step(plan)
    n_blur = 1 + 5 * plan.octave_max
    achieved_all = bbytes * args.steps / (blur_ms / 1e3) / 1e9 if blur_ms > 0 else None
    b0bytes = 5 * 12 * SIZE * SIZE
    achieved0 = b0bytes * args.steps / (blur0_ms / 1e3) / 1e9 if blur0_ms > 0 else None
    achieved_first = 8 * SIZE * SIZE * args.steps / (first_ms / 1e3) / 1e9 if first_ms > 0 else None
    line = {
        "metric": "keypoints/sec on 4096x4096 float32", "value": nkp / dev_s, "unit": "keypoints/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(world),
        "keypoints_per_image": nkp / args.steps / world,
        "timed_regions": {"repeats": len(regions), "reported": "median", "ms_per_step_each":
                          [1e3 * x / args.steps for x in all_dev_s]},
        "e2e": {"value": e2e_kp / e2e_s, "unit": "keypoints/s", "h2d_bytes_per_step": SIZE * SIZE * 4,
                "d2h_bytes_per_step": int(d2h / args.steps), "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "SiftPlan.keypoints_many (three images in flight; pinned host image in, numpy recarray out)",
                "one_at_a_time": {"api": "SiftPlan.keypoints(host image), the reference's call", "value":
                                  sync_kp / sync_s, "ms_per_step": 1e3 * sync_s / args.steps}},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm",
                     "kernel": "k_blur_tma, the whole Gaussian pyramid: all %d blur launches of a step "
                               "(normalise + first blur, then 5 blur+DoG per octave on %s planes)"
                               % (n_blur, " / ".join("%dx%d" % (int(w), int(h)) for w, h in plan.scales)),
                     "achieved": achieved_all, "peak": peak, "unit": "GB/s",
                     "frac": (achieved_all / peak) if achieved_all else None,
                     "traffic": (TRAFFIC_PER_STEP / n_blur) if TRAFFIC_PER_STEP else None, "traffic_note": TRAFFIC_NOTE,
                     "peak_source": peak_src + " (burst copy figure)",
                     "algorithmic_bytes_per_launch": bbytes / n_blur, "launches_per_step": n_blur,
                     "avg_launch_ms": blur_ms / args.steps / n_blur,
                     "algorithmic_bytes_per_step": bbytes, "ms_per_step": blur_ms / args.steps,
                     "octave0_launches": {"note": "the 5 blur+DoG launches on the 4096x4096 planes, 12*W*H bytes each",
                                          "achieved": achieved0, "frac": (achieved0 / peak) if achieved0 else None,
                                          "avg_launch_ms": blur0_ms / args.steps / 5},
                     "first_blur": {"note": "normalise fused into the first blur, 8*W*H bytes", "achieved": achieved_first,
                                    "frac": (achieved_first / peak) if achieved_first else None,
                                    "avg_launch_ms": first_ms / args.steps}},
        "stage_ms_per_step": {k: v / args.steps for k, v in sorted(stage_ms.items())},
        "stage_ms_note": "CUDA events on the plan's stream in a separate K-step region with profiling on",
        "wall_ms_per_step": 1e3 * wall / args.steps,
        "clocks": sampler.summary(),
    }
    if world > 1:
        line["with_gather"] = {
            "value": nkp / wg_s, "ms_per_step": 1e3 * wg_s / args.steps,
            "e2e_value": e2e_kp / wg_e2e_s, "e2e_ms_per_step": 1e3 * wg_e2e_s / args.steps,
            "exchange_host_ms_per_step": 1e3 * host_x[0] / (args.steps * len(withgather)),
            "note": "the same regions with every step's keypoint records all-gathered to every rank over NCCL "
                    "(dist.RecordExchange: one collective per step on a side stream, completed one pipeline step "
                    "later); not part of `value`: the images are independent, the path itself has no exchange step"}
    if not args.no_cpu_baseline and world == 1:
        from oracle import siftref
        img = np.array(host_imgs[0])
        siftref.set_num_threads(os.cpu_count())
        cores = siftref.num_threads()
        siftref.keypoints(img, octave_max=OCTAVES)  # warm-up
        v_all, ms_all = _cpu_leg(siftref, img, cores, 3)
        v_one, ms_one = _cpu_leg(siftref, img, 1, 1)
        line["cpu_baseline"] = {"value": v_all, "unit": "keypoints/s", "cores": cores, "kind": "port",
                                "ms_per_image": ms_all,
                                "single_thread": {"value": v_one, "ms_per_image": ms_one, "cores": 1},
                                "sample": "3 x one %dx%d image (seed 1234) on all %d host threads + 1 x the same image "
                                          "on one thread, oracle/libsiftref.so (OpenMP)" % (SIZE, SIZE, cores)}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
