"""BASELINE.json configs 3, 4 and 5 at their stated sizes, the LinearAlign options, device-resident data flow
and buffer overflow -- CUDA path (through the C ABI) against the oracle."""
import numpy as np
import pytest

from helpers import compare_whole, desc_sets, ms, same_records, sort_kp, sort_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sift():
    import sift_pyocl_b200
    return sift_pyocl_b200


# ---- config 3: the unit of work of "64 x 2048x2048 over 8 GPUs" is one 2048x2048 image -------------------
def test_config3_unit_2048(sift, oracle):
    plan, kp, ref = compare_whole(sift, oracle, ms(2048, 1234 + 17))
    assert plan.octave_max == 8 and kp.size > 10000
    # a second image through the same plan (plan reuse across the rank's share of the batch)
    img2 = ms(2048, 1234 + 18)
    assert same_records(plan.keypoints(img2), oracle.keypoints(img2))


# ---- config 4: MatchPlan 100k x 100k ----------------------------------------------------------------------
def test_config4_match_100k(sift, oracle):
    n = 100000
    k1, k2, perm = desc_sets(n, n, seed=7)
    mp = sift.MatchPlan()
    raw = mp.match(k1, k2, raw_results=True)
    assert raw.dtype == np.int32 and raw.shape[1] == 2 and mp.kpsize == n
    assert len(np.unique(raw[:, 0])) == len(raw)                       # at most one match per query
    # 4000 sampled queries (plus the first / last rows and the rows around the 2^16 boundaries) against the oracle
    rng = np.random.default_rng(0)
    sample = np.unique(np.concatenate([rng.choice(n, 4000, replace=False), np.arange(64), np.arange(n - 64, n),
                                       np.arange(65536 - 32, 65536 + 32)]))
    want = oracle.match(k1[sample], k2)
    want[:, 0] = sample[want[:, 0]]
    got = raw[np.isin(raw[:, 0], sample)]
    assert np.array_equal(sort_rows(got), sort_rows(want)) and len(want) > 3000
    # size-independent properties: every emitted pair points at the planted partner (list 2 row j is a perturbed
    # copy of list 1 row perm[j]); the same queries matched alone give the same pairs (queries are independent)
    assert np.array_equal(perm[raw[:, 1]], raw[:, 0])
    sub = mp.match(k1[sample], k2, raw_results=True)
    sub[:, 0] = sample[sub[:, 0]]
    assert np.array_equal(sort_rows(sub), sort_rows(want))
    # the (m, 2) recarray form is gathered on the device: same rows as host-side fancy indexing (match.py:267-270)
    res = mp.match(k1, k2)
    assert res.shape == (len(raw), 2) and res.dtype == mp.dtype_kp
    order = np.argsort(res[:, 0].x)
    by_q = raw[np.argsort(raw[:, 0])]
    assert np.array_equal(res[order, 0].desc, k1.desc[by_q[:, 0]]) and np.array_equal(res[order, 1].desc, k2.desc[by_q[:, 1]])
    assert np.array_equal(res[order, 1].x, k2.x[by_q[:, 1]])


# ---- config 5: LinearAlign on an 8192 x 8192 pair -----------------------------------------------------------
def test_config5_align_8192(sift, oracle):
    from scipy.ndimage import affine_transform
    ref = ms(8192, 1234)
    M = np.array([[1.01, -0.01], [0.005, 0.99]])
    off = np.array([7.0, 5.0])
    moved = affine_transform(ref, M, offset=off, order=1, mode="reflect").astype(np.float32)
    la = sift.LinearAlign(ref)
    assert la.sift.octave_max == 10
    # SiftPlan.keypoints at 8192^2: per-octave counts and the sorted records equal the oracle's
    want_kp, info = oracle.keypoints(ref, return_all=True)
    assert np.array_equal(la.sift.last_counts, info["n_per_octave"])
    assert same_records(la.ref_kp, want_kp) and la.ref_kp.size > 200000
    out = la.align(moved, return_all=True)
    assert out is not None and out["matching"].shape[0] > 50000
    assert same_records(out["keypoint"], oracle.keypoints(moved))
    # matched pairs == oracle.match, checked on 3000 sampled reference keypoints (the full 255k x 280k scan takes the
    # host cores half a minute): identical (reference index, frame index) pairs
    raw = la.match.last_pairs(raw_results=True)
    assert len(raw) == out["matching"].shape[0] and len(np.unique(raw[:, 0])) == len(raw)
    sample = np.unique(np.random.default_rng(1).choice(la.ref_kp.size, 3000, replace=False))
    raw_want = oracle.match(la.ref_kp[sample], out["keypoint"])
    raw_want[:, 0] = sample[raw_want[:, 0]]
    assert np.array_equal(sort_rows(raw[np.isin(raw[:, 0], sample)]), sort_rows(raw_want)) and len(raw_want) > 500
    by_q = raw[np.argsort(raw[:, 0])]
    order = np.lexsort((out["matching"][:, 0].angle, out["matching"][:, 0].scale, out["matching"][:, 0].y, out["matching"][:, 0].x))
    assert same_records(out["matching"][:, 0], la.ref_kp[by_q[:, 0]]) and same_records(out["matching"][:, 1], out["keypoint"][by_q[:, 1]])
    del order
    # scipy maps output -> input (moved[o] = ref[M o + off]); LinearAlign fits reference -> frame: the inverse map
    Minv = np.linalg.inv(M)
    assert np.allclose(out["matrix"], Minv, atol=2e-4) and np.allclose(out["offset"], -Minv.dot(off), atol=0.1)
    assert out["rms"] < 0.5
    want = oracle.transform(moved, out["matrix"], out["offset"], la.sift.buffers["min"].get()[0])
    assert np.array_equal(out["result"], want)
    core = (slice(256, -256), slice(256, -256))
    assert abs(out["result"] - ref)[core].mean() < 0.35 * abs(moved - ref)[core].mean()


# ---- LinearAlign options (alignment.py:149-154, 266-322): each against a numpy recomputation from the oracle ----
def _expected_pairs(oracle, ref_kp, kp):
    raw = oracle.match(ref_kp, kp)
    m = np.zeros((len(raw), 2), ref_kp.dtype).view(np.recarray)
    m[:, 0], m[:, 1] = ref_kp[raw[:, 0]], kp[raw[:, 1]]
    from sift_pyocl_b200.alignment import pairs_from_matching
    p = pairs_from_matching(m)
    return p[np.lexsort((p[:, 5], p[:, 4], p[:, 1], p[:, 0]))]  # canonical order: the fit does not depend on it


def _pair(seed=31, n=512):
    from scipy.ndimage import affine_transform
    ref = ms(n, seed)
    M = np.array([[1.02, -0.03], [0.02, 0.97]])
    off = np.array([4.0, -3.0])
    moved = affine_transform(ref, M, offset=off, order=1, mode="reflect").astype(np.float32)
    return ref, moved, M, off


def test_align_shift_only(sift, oracle):
    from sift_pyocl_b200.alignment import median_shift
    ref = ms(512, 41)
    moved = np.roll(ref, (6, -9), axis=(0, 1))
    la = sift.LinearAlign(ref)
    out = la.align(moved, shift_only=True, return_all=True)
    want_m, want_o = median_shift(_expected_pairs(oracle, oracle.keypoints(ref), oracle.keypoints(moved)))
    assert np.array_equal(out["matrix"], want_m) and np.array_equal(out["offset"], want_o)
    assert np.allclose(out["offset"], [6, -9], atol=0.05)
    assert np.array_equal(out["result"], oracle.transform(moved, want_m, want_o, la.sift.buffers["min"].get()[0]))
    assert np.array_equal(out["result"][64:-64, 64:-64], ref[64:-64, 64:-64])  # integer shift: exact realignment


def test_align_double_check_rejects_outliers(sift, oracle):
    from sift_pyocl_b200.alignment import affine_from_pairs, inlier_mask
    ref, moved, M, off = _pair(43)
    moved = moved.copy()
    # injected outliers: a patch of the frame is replaced by content from elsewhere in the frame, so its keypoints
    # match reference keypoints 250 px away
    moved[40:150, 40:150] = moved[300:410, 290:400].copy()
    la = sift.LinearAlign(ref)
    plain = la.align(moved, return_all=True)
    checked = la.align(moved, return_all=True, double_check=True)
    pairs = _expected_pairs(oracle, oracle.keypoints(ref), oracle.keypoints(moved))
    keep = inlier_mask(pairs)
    assert 0 < (~keep).sum() < len(keep) // 4
    want_plain = affine_from_pairs(pairs)
    want_checked = affine_from_pairs(pairs[keep])
    for got, want in ((plain, want_plain), (checked, want_checked)):
        assert np.allclose(got["matrix"], want[0], atol=1e-5) and np.allclose(got["offset"], want[1], atol=1e-3)
    truth = np.linalg.inv(M)  # scipy's matrix maps output -> input; LinearAlign fits reference -> frame
    assert abs(checked["matrix"] - truth).max() < abs(plain["matrix"] - truth).max()
    assert np.array_equal(checked["result"], oracle.transform(moved, checked["matrix"], checked["offset"],
                                                              la.sift.buffers["min"].get()[0]))


def test_align_relative_two_frames(sift, oracle):
    from scipy.ndimage import affine_transform
    from sift_pyocl_b200.alignment import affine_from_pairs, chain_transform
    ref, frame1, M, off = _pair(47)
    frame2 = affine_transform(frame1, np.array([[0.99, 0.02], [-0.01, 1.01]]), offset=[-2.0, 3.0], order=1,
                              mode="reflect").astype(np.float32)
    la = sift.LinearAlign(ref)
    o1 = la.align(frame1, return_all=True, relative=True)
    assert same_records(la.ref_kp, o1["keypoint"])          # the frame became the reference (alignment.py:304)
    o2 = la.align(frame2, return_all=True, relative=True)
    k0, k1, k2 = oracle.keypoints(ref), oracle.keypoints(frame1), oracle.keypoints(frame2)
    m1, f1 = affine_from_pairs(_expected_pairs(oracle, k0, k1))
    m2, f2 = affine_from_pairs(_expected_pairs(oracle, k1, k2))
    t1 = chain_transform(None, m1, f1)
    t2 = chain_transform(t1, m2, f2)
    assert np.allclose(o1["matrix"], t1[:2, :2], atol=1e-5) and np.allclose(o1["offset"], t1[:2, 2], atol=1e-3)
    assert np.allclose(o2["matrix"], t2[:2, :2], atol=1e-5) and np.allclose(o2["offset"], t2[:2, 2], atol=2e-3)
    assert np.allclose(la.relative_transfo, t2, atol=2e-3)
    # frame 2 mapped through the accumulated transform lands on the original reference
    core = (slice(80, -80), slice(80, -80))
    assert abs(o2["result"] - ref)[core].mean() < 0.5 * abs(frame2 - ref)[core].mean()


def test_align_roi_and_extra(sift, oracle):
    from sift_pyocl_b200.alignment import affine_from_pairs
    ref, moved, M, off = _pair(53)
    roi = np.zeros(ref.shape, np.int8)
    roi[100:400, 120:460] = 1
    la = sift.LinearAlign(ref, ROI=roi, extra=(8, 16))
    k0 = oracle.keypoints(ref)
    inside = roi[(np.round(k0.y).astype(np.int32), np.round(k0.x).astype(np.int32))].astype(bool)   # alignment.py:150-153
    assert same_records(la.ref_kp, k0[inside]) and 0 < inside.sum() < k0.size
    assert la.outshape == (512 + 16, 512 + 32)
    out = la.align(moved, return_all=True)
    want_m, want_o = affine_from_pairs(_expected_pairs(oracle, k0[inside], oracle.keypoints(moved)))
    assert np.allclose(out["matrix"], want_m, atol=1e-5) and np.allclose(out["offset"], want_o, atol=1e-3)
    assert out["result"].shape == (528, 544)
    want = oracle.transform(moved, out["matrix"], out["offset"], la.sift.buffers["min"].get()[0], (528, 544))
    assert np.array_equal(out["result"], want)
    # every matched reference keypoint lies inside the ROI
    mk = out["matching"][:, 0]
    assert roi[(np.round(mk.y).astype(np.int32), np.round(mk.x).astype(np.int32))].all()


def test_align_few_matches_falls_back_to_shift(sift, oracle):
    """Fewer than 18 matches -> translation only (alignment.py:266)."""
    ref = ms(96, 59)
    moved = np.roll(ref, (1, 2), axis=(0, 1))
    la = sift.LinearAlign(ref)
    out = la.align(moved, return_all=True)
    if out is not None and out["matching"].shape[0] < 18:
        assert np.array_equal(out["matrix"], np.identity(2, dtype=np.float32))


# ---- device-resident data flow ----------------------------------------------------------------------------
def test_match_device_resident_lists(sift, oracle):
    import torch
    from sift_pyocl_b200.match import DeviceRecords
    img = ms(512, 21)
    shifted = np.roll(img, (3, 5), axis=(0, 1))
    plan = sift.SiftPlan(template=img)
    ka = plan.keypoints(img)
    plan.submit(shifted)
    nb = plan.collect(records=False)            # records of the second image stay in HBM
    dev_b = plan.device_keypoints()
    assert isinstance(dev_b, DeviceRecords) and dev_b.size == nb and dev_b.shape == (nb,)
    kb = plan.fetch_keypoints()
    assert same_records(kb, oracle.keypoints(shifted)) and same_records(dev_b.get(), kb)
    mp = sift.MatchPlan()
    want = sort_rows(oracle.match(ka, kb))
    assert np.array_equal(sort_rows(mp.match(ka, dev_b, raw_results=True)), want)
    ta = torch.from_numpy(ka.view(np.uint8).reshape(-1, 144).copy()).cuda()   # torch tensor in place of pyopencl.array
    assert np.array_equal(sort_rows(mp.match(ta, dev_b, raw_results=True)), want)
    # hold(): list 0 stays resident, later calls with the same object skip the upload
    mp.hold(0, ka)
    coords = mp.match_coords(ka, dev_b)
    assert coords.shape == (len(want), 8)
    exp = np.stack([ka.x[want[:, 0]], ka.y[want[:, 0]], ka.scale[want[:, 0]], ka.angle[want[:, 0]],
                    kb.x[want[:, 1]], kb.y[want[:, 1]], kb.scale[want[:, 1]], kb.angle[want[:, 1]]], 1)
    assert np.array_equal(sort_rows(coords), sort_rows(exp))
    assert np.array_equal(sort_rows(mp.last_pairs(raw_results=True)), want)


def test_device_input_is_ordered_after_its_producer(sift, oracle):
    """A device-resident image still being written by an asynchronous kernel on torch's stream when keypoints() is
    called: the plan's private stream must wait for the producer."""
    import torch
    img = ms(1024, 61)
    want = oracle.keypoints(img)
    plan = sift.SiftPlan(shape=img.shape, dtype=np.float32)
    src = torch.from_numpy(img).cuda()
    big = torch.empty((64, 1024, 1024), device="cuda")
    for _ in range(3):
        buf = torch.zeros_like(src)
        torch.cuda.synchronize()
        for i in range(20):           # ~ms of queued work ahead of the producer of `buf`
            big.mul_(1.0001)
        buf.copy_(src)                 # asynchronous on torch's current stream
        kp = plan.keypoints(buf)
        assert same_records(kp, want)
    with pytest.raises(AssertionError):  # RGB plan fed with float32 (H, W, 3) data
        sift.SiftPlan(shape=(64, 64, 3), dtype=np.uint8).keypoints(np.zeros((64, 64, 3), np.float32))


def test_keypoints_many_drains_when_abandoned(sift, oracle):
    imgs = [ms(256, 70 + i) for i in range(5)]
    plan = sift.SiftPlan(shape=imgs[0].shape, dtype=np.float32)
    gen = plan.keypoints_many(imgs)
    first = next(gen)
    gen.close()                        # consumer stops early with two images still in flight
    assert same_records(first, oracle.keypoints(imgs[0]))
    assert same_records(plan.keypoints(imgs[3]), oracle.keypoints(imgs[3]))   # the plan is usable again

    def boom():
        yield imgs[0]
        yield imgs[1]
        raise ValueError("source failed")
    with pytest.raises(ValueError):
        list(plan.keypoints_many(boom()))
    assert same_records(plan.keypoints(imgs[4]), oracle.keypoints(imgs[4]))


def test_two_lanes_images_processed_concurrently(sift, oracle, monkeypatch):
    """Images in flight together alternate between the plan's two compute streams (each with its own planes and
    lists); every result must equal the one-image-at-a-time result, for converted (uint8) input too."""
    rng = np.random.default_rng(5)
    imgs = [ms(1024, 90 + i) for i in range(6)]
    lo, hi = min(i.min() for i in imgs), max(i.max() for i in imgs)
    imgs8 = [((i - lo) / (hi - lo) * 255).astype(np.uint8) for i in imgs]
    want = [oracle.keypoints(i) for i in imgs8]
    plan = sift.SiftPlan(shape=imgs8[0].shape, dtype=np.uint8)
    one_lane_bytes = plan.memory
    for _ in range(2):       # second round: both lanes exist already
        got = list(plan.keypoints_many(imgs8))
        assert all(same_records(g, w) for g, w in zip(got, want))
    assert plan.memory > 1.5 * one_lane_bytes            # the second lane was allocated on demand
    assert same_records(plan.keypoints(imgs8[2]), want[2])       # and the one-image path still works
    del rng
    # SIFTB_LANES=1: every image on one compute stream, no second set of planes
    monkeypatch.setenv("SIFTB_LANES", "1")
    single = sift.SiftPlan(shape=imgs8[0].shape, dtype=np.uint8)
    got = list(single.keypoints_many(imgs8[:4]))
    assert all(same_records(g, w) for g, w in zip(got, want))
    assert single.memory == one_lane_bytes


# ---- keypoint buffer overflow: clean truncation (reference only warns, plan.py:771) ------------------------
def test_overflow_truncates_cleanly(sift, oracle):
    img = ms(512, 81)
    full = sift.SiftPlan(template=img).keypoints(img)
    small = sift.SiftPlan(template=img, PIX_PER_KP=400)    # kpsize = 655 slots per octave, 1310 rows in the list
    assert small.kpsize == 512 * 512 // 400
    kp = small.keypoints(img)
    assert 0 < kp.size <= 2 * small.kpsize and kp.size == small.last_counts.sum()
    # every returned record is a genuine keypoint of the image (which ones survive depends on the atomics' order)
    key = lambda k: set(map(bytes, np.ascontiguousarray(k).view(np.uint8).reshape(-1, 144)))  # noqa: E731
    assert key(kp) <= key(full) and len(key(kp)) == kp.size
    assert same_records(small.keypoints(ms(0, 82, (512, 512)) * 0 + 1.0), full[:0])   # plan still healthy: flat image
    med = sift.SiftPlan(template=img, PIX_PER_KP=150)     # overflow only in the extra-orientation rows / octave 0
    kp2 = med.keypoints(img)
    assert key(kp2) <= key(full) and len(key(kp2)) == kp2.size and kp2.size == med.last_counts.sum()


# ---- profile=True on MatchPlan and LinearAlign (match.py:226-263, alignment.py:363-376) ---------------------
def test_profile_events_match_and_align(sift, capsys):
    k1, k2, _ = desc_sets(2000, 1500)
    mp = sift.MatchPlan(profile=True)
    mp.match(k1, k2)
    names = [n for n, _ in mp.events]
    assert "matching" in names and any("KP_1" in n for n in names) and all(ms_ >= 0 for _, ms_ in mp.events)
    mp.log_profile()
    assert "matching" in capsys.readouterr().out
    mp.reset_timer()
    assert mp.events == []
    ref, moved, _, _ = _pair(31, 384)
    la = sift.LinearAlign(ref, profile=True)
    assert la.align(moved) is not None
    assert any(n.startswith("transform") for n, _ in la.events) and any(n == "matching" for n, _ in la.match.events)
    la.log_profile()
    out = capsys.readouterr().out
    assert "transform" in out and "matching" in out and "descriptors" in out


# ---- devicetype="GPU": the numbers of orientation_gpu.cl + keypoints_gpu2.cl ---------------
def _stage_inputs(oracle, shape=(240, 320), seed=9, octsize=1):
    g0 = oracle.blur(oracle.normalize(ms(0, seed, shape)), oracle.gaussian_taps(1.5199))
    G, D = oracle.pyramid_octave(g0)
    out = []
    for s in (1, 2, 3):
        ko, no = oracle.local_maxmin(D, s, octsize=octsize)
        kc, nc = oracle.compact(oracle.interp_keypoint(D, ko, 0, no), 0, no)
        grad, ori = oracle.gradient(G[s])
        out.append((kc, nc, grad, ori))
    return out


@pytest.mark.parametrize("octsize", [1, 4])
def test_gpu_variant_stages(oracle, octsize):
    from sift_pyocl_b200 import stages
    differs = 0
    for kc, nc, grad, ori in _stage_inputs(oracle, octsize=octsize):
        want, nw = oracle.orientation(kc, grad, ori, 0, nc, octsize=octsize, variant="gpu")
        got, ng = stages.orientation(kc[:nc], grad, ori, octsize, variant="gpu")
        assert ng == nw and nw >= nc
        assert np.array_equal(got[:nc], want[:nc])
        assert np.array_equal(sort_rows(got[nc:]), sort_rows(want[nc:nw]))
        cpu, _ = oracle.orientation(kc, grad, ori, 0, nc, octsize=octsize)
        differs += int((cpu[:nc, 3] != want[:nc, 3]).sum())          # the two families really differ ...
        assert np.median(abs(cpu[:nc, 3] - want[:nc, 3])) < 2e-2      # ... mostly by the bin offset and the wrap
        dw = oracle.descriptor(want, grad, ori, 0, nw, octsize=octsize, variant="gpu")[:nw]
        dg = stages.descriptor(want[:nw], grad, ori, octsize, variant="gpu")
        assert dw.any() and np.array_equal(dg, dw)
        dc = oracle.descriptor(want, grad, ori, 0, nw, octsize=octsize)[:nw]
        assert (dc != dw).any() and np.median(np.abs(dc.astype(int) - dw.astype(int))) <= 1   # 1e-5 quantisation
    assert differs > 0


def test_gpu_variant_whole_path(sift, oracle):
    img = ms(512, 1234)
    plan = sift.SiftPlan(template=img, devicetype="GPU")
    assert plan.variant == "gpu"
    kp = plan.keypoints(img)
    want, info = oracle.keypoints(img, return_all=True, variant="gpu")
    assert np.array_equal(plan.last_counts, info["n_per_octave"]) and same_records(kp, want)
    cpu = sift.SiftPlan(template=img).keypoints(img)       # default devicetype: the *_cpu.cl numbers
    assert same_records(cpu, oracle.keypoints(img)) and not same_records(cpu, kp)
    # keypoints with enormous windows (custom init_sigma): the fixed [-64, 64) window of keypoints_gpu2.cl truncates them
    big = sift.SiftPlan(template=img, devicetype="GPU", init_sigma=3.5)
    assert same_records(big.keypoints(img), oracle.keypoints(img, init_sigma=3.5, variant="gpu"))


@pytest.mark.parametrize("init_sigma", [2.5, 3.5])
def test_large_init_sigma_windows_beyond_the_row_table(sift, oracle, init_sigma):
    """Descriptor windows with more rows than the per-keypoint interval table holds (iradius >= 52) take the
    full-scan path; warps mix tabled, full-scan and idle octets (a divergent warp barrier there hung round 1's
    kernel for init_sigma >= 2.5)."""
    img = ms(384, 91)
    plan, kp, ref = compare_whole(sift, oracle, img, init_sigma=init_sigma)
    assert kp.size > 50 and kp.scale.max() > 3 * init_sigma


def test_match_l2_metric_option(sift, oracle):
    """The optional squared-L2 matcher (not the reference's metric) against its own oracle, short and long lists."""
    k1, k2, _ = desc_sets(3000, 2500, seed=5)
    mp = sift.MatchPlan()
    l1 = mp.match(k1, k2, raw_results=True)
    mp.metric = "l2"
    got = mp.match(k1, k2, raw_results=True)
    assert np.array_equal(sort_rows(got), sort_rows(oracle.match(k1, k2, metric="l2"))) and len(got) > 2000
    k1, k2, _ = desc_sets(80000, 40000, seed=6)       # two queries per thread, segmented second list
    got = mp.match(k1, k2, raw_results=True)
    sample = np.arange(0, 80000, 37)
    want = oracle.match(k1[sample], k2, metric="l2")
    want[:, 0] = sample[want[:, 0]]
    assert np.array_equal(sort_rows(got[np.isin(got[:, 0], sample)]), sort_rows(want)) and len(want) > 500
    mp.metric = "l1"
    assert np.array_equal(sort_rows(mp.match(k1[:3000], k2[:2500], raw_results=True)),
                          sort_rows(oracle.match(k1[:3000], k2[:2500])))
    assert len(l1) > 2000
