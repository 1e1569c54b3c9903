#!/usr/bin/env python
"""bench.py -- keypoints/sec of SiftPlan.keypoints() on 4096x4096 float32 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # the CPU arm (oracle port, all host threads)

A "step" is one pass of the hot path (plan.keypoints) over one synthetic image per GPU.  Workload =
BASELINE.json configs[1]: SiftPlan 4096x4096 float32, 3 octaves x 3 scales (par.OctaveMax = 3), images
= seeded "multiscale noise" (sift_pyocl_b200.utils.multiscale_image, seed 1234 + index).
  value : whole-job keypoints/s with the images already resident in HBM (device pointer input), records
          left on the device, CUDA events on the plan's stream, max over ranks;
  e2e   : the same metric through the public API with HOST buffers: pinned host image -> H2D copy ->
          kernels -> D2H copy of the records -> numpy recarray, every step;
  roofline : the Gaussian blur(+DoG) kernel family, algorithmic bytes (12*W*H per blur+DoG launch,
          8*W*H for the first blur) / CUDA-event time of those launches inside the timed
          region, against MEASURED_PEAKS.json hbm_gbs;
  cpu_baseline : the oracle (CPU port of the reference kernels, OpenMP) on the same image.
For N > 1 (torchrun) every rank runs the same per-GPU work on different images (weak scaling) and
each step ends with the NCCL all-gather of the keypoint records (counts + padded payload).
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
# dram__bytes_read.sum + dram__bytes_write.sum per octave-0 blur+DoG launch from the committed ncu --set full
# capture (profiles/r01c_ncu_full_final.csv), averaged over the five tap counts; None until measured
TRAFFIC_PER_LAUNCH = 164.3e6


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
            "config": {"workload": _workload_name(), "image": "multiscale noise seed 1234"},
            "cpu_baseline": {"value": v, "unit": "keypoints/s", "cores": cores, "kind": "port",
                             "sample": "%d x one %dx%d image, OpenMP on %d threads" % (args.steps, SIZE, SIZE, cores)},
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
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

    def gather(n):
        if world > 1:
            sdist.allgather_records(sdist.device_records_tensor(plan, n))

    def device_step(i):
        plan.submit(dev_imgs[i % N_IMAGES])
        n = plan.collect(records=False)
        gather(n)
        return n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        device_step(i)
    plan.set_profile(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = plan.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    nkp, blur_ms, blur0_ms, stage_ms = 0, 0.0, 0.0, {}
    ev0.record(stream)
    t0 = time.perf_counter()
    # K steps, software-pipelined: the kernels of step i+1 are enqueued before step i is collected, so the
    # all-gather of step i's records (N > 1) and the host-side bookkeeping overlap the next image's kernels
    plan.submit(dev_imgs[0])
    for i in range(args.steps):
        if i + 1 < args.steps:
            plan.submit(dev_imgs[(i + 1) % N_IMAGES])
        n = plan.collect(records=False)
        gather(n)
        nkp += n
        for name, ms in plan.fetch_events():
            key = name.split(" octave")[0]
            stage_ms[key] = stage_ms.get(key, 0.0) + ms
            if "blur" in name:
                blur_ms += ms
            if name == "blur + DoG octave 0":
                blur0_ms += ms
    cur = torch.cuda.current_stream()
    cur.wait_stream(stream)  # the gather (if any) runs on torch's stream after the plan's stream
    ev1.record(cur)
    barrier()
    wall = time.perf_counter() - t0
    dev_s = ev0.elapsed_time(ev1) / 1e3
    launches = plan.launches - launches0
    plan.set_profile(False)

    # end to end through the public API with host buffers: every step copies its image from pinned host
    # memory to the device and its records back into host memory.  SiftPlan.keypoints_many keeps three images in
    # flight so the copies of one image overlap the kernels of the others (all of it inside the timed region).
    for kp in plan.keypoints_many(host_imgs[i % N_IMAGES] for i in range(3)):
        pass
    barrier()
    e2e_kp, d2h, t0 = 0, 0, time.perf_counter()
    pending = None  # N > 1: the all-gather of step i is completed while step i+1 runs (counts first, then payload)
    for kp in plan.keypoints_many(host_imgs[i % N_IMAGES] for i in range(args.steps)):
        e2e_kp += kp.size
        d2h += kp.size * 144 + 4 * (1 + 13 * plan.octave_max + 2)
        if world > 1:
            started = sdist.allgather_records_begin(sdist.device_records_tensor(plan, kp.size))
            if pending is not None:
                pending.finish()
            pending = started
    if pending is not None:
        pending.finish()
    barrier()
    e2e_s = time.perf_counter() - t0
    # the same, strictly one image at a time (SiftPlan.keypoints, the reference's call)
    t0 = time.perf_counter()
    for i in range(args.steps):
        plan.keypoints(host_imgs[i % N_IMAGES])
    barrier()
    e2e_sync_s = time.perf_counter() - t0
    sampler.stop_flag = True

    if world > 1:
        t = torch.tensor([dev_s, e2e_s, float(nkp), float(e2e_kp), float(launches)], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dev_s, e2e_s = float(tmax[0]), float(tmax[1])
        nkp, e2e_kp, launches = float(t[2]), float(t[3]), int(t[4])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = _peaks()
    bbytes = blur_bytes_per_step(plan) * args.steps
    achieved_all = bbytes / (blur_ms / 1e3) / 1e9 if blur_ms > 0 else None
    # dominant kernel: the five blur+DoG launches on the full-resolution planes (octave 0), 12*W*H bytes each
    b0bytes = 5 * 12 * SIZE * SIZE
    achieved = b0bytes * args.steps / (blur0_ms / 1e3) / 1e9 if blur0_ms > 0 else None
    line = {
        "metric": "keypoints/sec on 4096x4096 float32", "value": nkp / dev_s, "unit": "keypoints/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": _workload_name(), "images_per_step_per_gpu": 1,
                   "image": "multiscale noise, seeds 1234+", "keypoints_per_image": nkp / args.steps / world,
                   "l2": "inputs larger than L2: %d distinct 67 MB images cycled, ~1.2 GB of planes rewritten per step"
                         % N_IMAGES,
                   "gather": "NCCL all-gather of records per step" if world > 1 else "none (1 GPU)"},
        "e2e": {"value": e2e_kp / e2e_s, "unit": "keypoints/s", "h2d_bytes_per_step": SIZE * SIZE * 4,
                "d2h_bytes_per_step": int(d2h / args.steps), "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "SiftPlan.keypoints_many (3 images in flight)",
                "ms_per_step_one_at_a_time": 1e3 * e2e_sync_s / args.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm",
                     "kernel": "k_blur_tma: the 5 blur+DoG launches on the 4096x4096 planes (octave 0), 11..27 taps",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": TRAFFIC_PER_LAUNCH, "peak_source": peak_src + " (burst copy figure)",
                     "algorithmic_bytes_per_launch": 12 * SIZE * SIZE, "launches_per_step": 5,
                     "avg_launch_ms": blur0_ms / args.steps / 5,
                     "all_blur_launches": {"note": "all %d blur launches of the step incl. first blur and octaves >= 1"
                                                   % (1 + 5 * plan.octave_max),
                                           "achieved": achieved_all, "frac": (achieved_all / peak) if achieved_all else None,
                                           "algorithmic_bytes_per_step": bbytes // args.steps,
                                           "ms_per_step": blur_ms / args.steps}},
        "stage_ms_per_step": {k: v / args.steps for k, v in sorted(stage_ms.items())},
        "wall_ms_per_step": 1e3 * wall / args.steps,
        "clocks": sampler.summary(),
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle import siftref
        siftref.set_num_threads(os.cpu_count())
        img = np.array(host_imgs[0])
        siftref.keypoints(img, octave_max=OCTAVES)
        reps, t0 = 3, time.perf_counter()
        for _ in range(reps):
            n_cpu = siftref.keypoints(img, octave_max=OCTAVES).size
        dt = (time.perf_counter() - t0) / reps
        line["cpu_baseline"] = {"value": n_cpu / dt, "unit": "keypoints/s", "cores": siftref.num_threads(),
                                "kind": "port", "ms_per_image": 1e3 * dt,
                                "sample": "%d x one %dx%d image (seed 1234), oracle/libsiftref.so OpenMP" % (reps, SIZE, SIZE)}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
