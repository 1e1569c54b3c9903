"""One pass over the BASELINE.json configs 2-5 on one GPU; prints one JSON line per config (not the bench contract,
just the numbers quoted in DESIGN.md / profiles)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch
import sift_pyocl_b200 as sift
from sift_pyocl_b200._lib import dtype_kp
from sift_pyocl_b200.utils import multiscale_image
from oracle import siftref


def sync():
    torch.cuda.synchronize()


def timed(fn, reps):
    fn()
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    sync()
    return (time.perf_counter() - t0) / reps, out


which = sys.argv[1:] or ["2", "3", "4", "5"]
if "2" in which:  # SiftPlan 4096^2, all 9 octaves (reference-faithful) and 3 octaves
    img = multiscale_image(4096, 1234)
    dimg = torch.from_numpy(img).cuda()
    for octs in (100000, 3):
        sift.par["OctaveMax"] = octs
        plan = sift.SiftPlan(template=img)
        sift.par["OctaveMax"] = 100000

        def run():
            plan.submit(dimg)
            return plan.collect(records=False)

        def run_stream(k=20):  # a stream of images, two in flight (the plan's two compute lanes)
            plan.submit(dimg)
            for i in range(k):
                if i + 1 < k:
                    plan.submit(dimg)
                plan.collect(records=False)
        dt, n = timed(run, 10)
        dts, _ = timed(run_stream, 2)
        print(json.dumps({"config": 2, "octaves": plan.octave_max, "ms_one_image": 1e3 * dt,
                          "ms_per_image_streamed": 1e3 * dts / 20, "keypoints": int(n),
                          "keypoints_per_s_streamed": n / (dts / 20), "per_octave": plan.last_counts.tolist()}))
        del plan
if "3" in which:  # 2048^2 images, this GPU's share of a 64-image batch (8 images)
    imgs = [multiscale_image(2048, 1234 + i) for i in range(8)]
    plan = sift.SiftPlan(template=imgs[0])
    pinned = []
    for im in imgs:
        b = plan.pinned_empty(im.shape)
        b[...] = im
        pinned.append(b)
    dt, kps = timed(lambda: list(plan.keypoints_many(pinned)), 3)
    nk = sum(k.size for k in kps)
    ref = siftref.keypoints(imgs[0])
    ok = np.array_equal(np.sort(kps[0].x), np.sort(ref.x)) and kps[0].size == ref.size
    print(json.dumps({"config": 3, "images": 8, "ms_per_image": 1e3 * dt / 8, "images_per_s": 8 / dt,
                      "keypoints_per_s": nk / dt, "keypoints_per_image": nk / 8, "matches_oracle_image0": bool(ok)}))
    del plan
if "4" in which:  # MatchPlan 100k x 100k (L1, ratio 0.73^2)
    rng = np.random.default_rng(3)
    n = 100000
    d1 = np.minimum(rng.gamma(1.0, 28.0, (n, 128)), 255).astype(np.uint8)
    perm = rng.permutation(n)
    d2 = np.clip(d1[perm].astype(np.int16) + rng.integers(-2, 3, (n, 128)) * (rng.random((n, 128)) < 0.5), 0, 255)
    k1, k2 = np.zeros(n, dtype_kp), np.zeros(n, dtype_kp)
    k1["desc"], k2["desc"] = d1, d2.astype(np.uint8)
    k1, k2 = k1.view(np.recarray), k2.view(np.recarray)
    mp = sift.MatchPlan(profile=True)
    dt, raw = timed(lambda: mp.match(k1, k2, raw_results=True), 3)
    mp.hold(0, k1)
    mp.hold(1, k2)
    dt_res, _ = timed(lambda: mp.match(k1, k2, raw_results=True), 3)   # both lists resident on the device
    kernel_ms = sorted(ms for name, ms in mp.events if name == "matching")[:-1]   # 8 runs, the slowest (first) dropped
    kernel_ms = kernel_ms[len(kernel_ms) // 2]
    sub = np.sort(rng.choice(n, 4000, replace=False))
    want = siftref.match(k1[sub], k2)
    got = raw[np.isin(raw[:, 0], sub)]
    got = got[np.argsort(got[:, 0])]
    ok = np.array_equal(np.searchsorted(sub, got[:, 0]), want[:, 0]) and np.array_equal(got[:, 1], want[:, 1])
    inv = np.empty(n, np.int64)
    inv[perm] = np.arange(n)
    # VABSDIFF4.U8.ACC issues at 64 lanes/clk/SM (tools/sad_probe.cu): the floor of a brute-force L1 scan
    sm_clock = torch.cuda.clock_rate() * 1e6 if hasattr(torch.cuda, "clock_rate") else 1.965e9
    floor_ms = 1e3 * (n * n * 32) / (64 * 148 * 1.965e9)
    print(json.dumps({"config": 4, "n1": n, "n2": n, "ms_host_lists": 1e3 * dt, "ms_resident_lists": 1e3 * dt_res,
                      "kernel_ms": kernel_ms, "vabsdiff4_floor_ms_at_1965MHz": floor_ms,
                      "kernel_frac_of_floor": floor_ms / kernel_ms, "matches": int(len(raw)),
                      "byte_sad_per_s_kernel": n * n * 128 / (kernel_ms / 1e3), "subset_4000_equals_oracle": bool(ok),
                      "true_pairs_found": int((inv[raw[:, 0]] == raw[:, 1]).sum())}))
if "5" in which:  # LinearAlign 8192^2 pair
    from scipy.ndimage import affine_transform
    ref = multiscale_image(8192, 77)
    M = np.array([[1.01, -0.02], [0.015, 0.99]])
    off = np.array([7.0, 5.0])
    moved = affine_transform(ref, M, offset=off, order=1, mode="reflect").astype(np.float32)
    t0 = time.perf_counter()
    la = sift.LinearAlign(ref, profile=True)
    t_init = time.perf_counter() - t0
    la.align(moved)  # warm-up (buffers of the matcher and of the warp are allocated on first use)
    la.sift.reset_timer(), la.match.reset_timer()
    la.events = []
    t0 = time.perf_counter()
    res = la.align(moved)                      # the reference's call: aligned image only
    t_align = time.perf_counter() - t0
    dev = {}
    for name, ms in la.sift.events + la.match.events:
        key = name.split(" octave")[0]
        dev[key] = dev.get(key, 0.0) + ms
    t0 = time.perf_counter()
    out = la.align(moved, return_all=True)     # + keypoints and matched records on the host
    t_all = time.perf_counter() - t0
    core = (slice(256, -256), slice(256, -256))
    print(json.dumps({"config": 5, "ref_keypoints": int(la.ref_kp.size), "init_s": t_init, "align_s": t_align,
                      "align_return_all_s": t_all, "device_ms": {k: round(v, 3) for k, v in sorted(dev.items())},
                      "device_ms_total": sum(dev.values()),
                      "matches": int(out["matching"].shape[0]), "rms": float(out["rms"]),
                      "err_before": float(abs(moved - ref)[core].mean()),
                      "err_after": float(abs(out["result"] - ref)[core].mean()),
                      "matrix": out["matrix"].tolist(), "offset": out["offset"].tolist()}))
