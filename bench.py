#!/usr/bin/env python
"""bench.py -- atom-swap proposals/s + morph frames/s at 1024^2 RGBA (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size 1024] [--frames 64]

Workload (BASELINE.json configs[1]): synthetic 1024x1024 RGBA, 2 key frames (full square -> disc),
1 048 576 atoms, spline motion + cosine fading, 64 output frames.

One STEP = one pass of the hot path over one batch:
    render phase : the 64 output frames of the morph (k_bin2 + k_acc per batch of 8 frames)
    swap phase   : SWAP_ROUNDS rounds of disjoint pair-swap proposals on the 1M-atom chain
Both phases are timed separately with CUDA events on the engine's stream, inputs resident in HBM.
`value` is the render throughput (frames/s); the swap throughput and its roofline are reported in
the `swap` object of the same JSON line.  `e2e` is the same frames/s measured through the C-ABI with
HOST buffers: per step the trajectory table is uploaded (H2D), the frames are rendered and copied
back (D2H) inside the timed region.

N > 1 (torchrun): output frames are sharded by frame range, swap rounds by atom range; the timed
region is bracketed by a barrier and the time is the max over ranks (weak scaling: every rank
renders `--frames` frames and proposes SWAP_ROUNDS rounds on its slice).
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
sys.path.insert(0, ROOT)

SWAP_EPOCHS = 4           # matcher steps per timed step; one matcher step = `world` re-tiled epochs of 64 rounds on the rank's part + 1 exchange
SWAP_ROUNDS = 64 * SWAP_EPOCHS
HBM_FALLBACK_GBS = 6650.0


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, BEFORE any pinned buffer is allocated: page-locked
    memory is placed on the node of the allocating thread, and a frame read back over PCIe into the other socket's
    memory crosses the inter-socket link.  Returns a short description for the bench line."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return {"numa_node": None, "note": "no NUMA information for %s" % bus}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception as ex:
        return {"numa_node": None, "note": "not bound: %r" % (ex,)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------ reference arm
def build_reference(size, seed_images, chains, blobs_per_frame, params):
    from oracle import amref
    m = amref.RefMorph(**params)
    for k, im in enumerate(seed_images):
        m.add_image(k, im)
    m.set_resolution(size, size)
    for k, blobs in enumerate(blobs_per_frame):
        m.import_blobs(k, blobs)
    for c in chains:
        m.import_chain(c["key"], c["words"], c["max_surface"])
    m.finish_import()
    return m


def reference_tables(size, images):
    """Chain table + blobs for the reference arm WITHOUT the GPU: the single-blob C2 scene has a closed
    form (row-major ranks; duplicates follow init_morph's rule with a numpy shuffle)."""
    rng = np.random.default_rng(4321)
    cols = []
    blobs = []
    W = 0
    for im in images:
        W = max(W, int((im[..., 3] != 0).sum()))
    for im in images:
        ys, xs = np.nonzero(im[..., 3] != 0)
        words = xs.astype(np.uint64) | (ys.astype(np.uint64) << np.uint64(16)) | (np.uint64(3) << np.uint64(48))
        n = len(words)
        if n < W:
            perm = rng.permutation(n)
            src = perm[np.arange(n, W) % n]
            fr = rng.integers(0, 256, size=(W - n, 2)).astype(np.uint64)
            dup = (words[src] & np.uint64(0xffffffff)) | (fr[:, 0] << np.uint64(32)) | (fr[:, 1] << np.uint64(40)) | (np.uint64(1) << np.uint64(48))
            words = np.concatenate([words, dup])
        cols.append(words)
        c = im[ys, xs].astype(np.float64) / 255.0
        blobs.append([dict(group=0, stats=np.array([xs.mean(), ys.mean(), c[:, 0].mean(), c[:, 1].mean(), c[:, 2].mean(), c[:, 3].mean()]),
                           surface=(ys.astype(np.uint64) * np.uint64(65536) + xs.astype(np.uint64)))])
    return [dict(key=0, words=np.stack(cols), max_surface=W)], blobs


def cpu_reference_numbers(m, frames, frame_budget_s=25.0, morph_steps=16, gpu_render=None):
    """Times the reference's own CPU path: get_pixels(t) on one thread (the reference renderer is
    single-threaded) and thread::morph() with all host threads.  The frames it renders are compared with the
    GPU's frames at the same times on the same chain table (gpu_render(t) -> packed RGBA image)."""
    from oracle import amref
    cores = amref.hardware_concurrency()
    # render: frames until the budget is used (at least one)
    t_used, n = 0.0, 0
    parity = {"frames": 0, "max_lsb": 0, "px_diff": 0, "pixels": 0}
    while n < frames and (n == 0 or t_used < frame_budget_s):
        # spread the sampled frames over the morph (0, 1/2, 1/4, 3/4, ... of it) instead of its first frames only
        k = [0, frames // 2, frames // 4, (3 * frames) // 4, frames // 8, (5 * frames) // 8][n] if n < 6 and frames >= 8 else n
        dt, ref = m.time_render(k / float(frames))
        t_used += dt
        n += 1
        if gpu_render is not None:
            got = gpu_render(k / float(frames))
            d = np.abs(amref.unpack_rgba(ref).astype(np.int16) - amref.unpack_rgba(got).astype(np.int16))
            parity["frames"] += 1
            parity["max_lsb"] = max(parity["max_lsb"], int(d.max()))
            parity["px_diff"] += int((d.max(axis=-1) != 0).sum())
            parity["pixels"] += int(ref.size)
    fps = n / t_used
    m.set(threads=cores, cycle_length=100000)
    m.sync()
    dt = m.time_morph_steps(morph_steps)
    pps = morph_steps * max(1, cores) * 100000 / dt
    m.sync()
    return dict(fps=fps, frames=n, render_s=t_used, pps=pps, cores=int(cores), morph_s=dt, parity=parity)


def other_configs(dev, peak):
    """BASELINE.json configs 3, 4 and 5 on one GPU, device-resident frames/s (a few seconds each; N = 1 only).  Bytes per frame
    are SURVEY.md section 8d's: 24 B (linear or <= 3 key frames) / 40 B (spline, >= 4 key frames) per atom + 4 B per pixel; the
    fluid configuration adds 184 B per particle and MPM step."""
    import time
    import torch
    from atomorph_b200 import engine as eng
    from atomorph_b200 import scenes
    res = {}

    def timed_frames(e, times, size, reps):
        out = torch.empty((len(times), size, size), dtype=torch.int32, device=dev)
        e.render_into(times, out.data_ptr(), True)
        e.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            e.render_into(times, out.data_ptr(), True)
        e.sync()
        dt = (time.perf_counter() - t0) / reps
        del out
        return dt

    def entry(name, what, n, dt, bytes_per_frame, e):
        res[name] = {"workload": what, "frames_per_s": n / dt, "us_per_frame": 1e6 * dt / n, "bytes_per_frame": int(bytes_per_frame),
                     "roofline_frac": bytes_per_frame * n / dt / 1e9 / peak, "path_frames": e.render_path_frames()}

    try:
        # C4: 512^2, 2 500 blobs per key frame (one chain per blob group), density 2
        e = eng.Engine(dev.index, seed=1, motion=eng.LINEAR, fading=eng.COSINE, density=2, blob_rgba_weight=2, blob_size_weight=1, blob_xy_weight=3, threads=0, cycle_length=1000)
        e.load_images(scenes.rect_blobs(512, 2500, frames=2, seed=11, min_side=2, max_side=20))
        e.blobify(); e.match_init(); e.match_rounds(2000); e.init_chains(); e.swap_rounds(400); e.render_prepare()
        A = e.table_device_ptr(0)[1]
        dt = timed_frames(e, np.array([f / 128.0 for f in range(128)]), 512, 3)
        entry("c4", "512x512, 2500 blobs per key frame, density 2, linear + cosine, 128 frames (several chains: general A-buffer path)", 128, dt, A * 24 + 512 * 512 * 4, e)
        del e
        # C5: 4096^2, 8 cyclic key frames, 16.7 M atoms, 64 of the 512 frames
        e = eng.Engine(dev.index, seed=1, motion=eng.SPLINE, fading=eng.COSINE, threads=0, cycle_length=1000)
        e.load_images(scenes.rotating_shapes(4096, 8))
        e.step(8); e.swap_rounds(256); e.render_prepare()
        A = e.table_device_ptr(0)[1]
        dt = timed_frames(e, np.array([f / 512.0 for f in range(64)]), 4096, 2)
        entry("c5", "4096x4096, 8 cyclic key frames, %d atoms, spline + cosine, 64 of 512 frames (--workload c5 is the multi-GPU run)" % A, 64, dt, A * 40 + 4096 * 4096 * 4, e)
        del e
        torch.cuda.empty_cache()
        # C3: 1024^2, fluid (10 MPM steps per frame) + perlin fading + feather 2
        e = eng.Engine(dev.index, seed=1, motion=eng.SPLINE, fading=eng.PERLIN, feather=2, fluid=10, threads=0, cycle_length=1000)
        e.load_images(scenes.square_to_disc(1024))
        e.step(8); e.swap_rounds(512); e.render_prepare()
        A = e.table_device_ptr(0)[1]
        dt = timed_frames(e, np.array([f / 64.0 for f in range(16)]), 1024, 1)
        entry("c3", "1024x1024, 2 key frames, fluid 10 MPM steps per frame + perlin + feather 2, 16 frames (stateful particle path)", 16, dt, A * 24 + 1024 * 1024 * 4 + 10 * 184 * A, e)
        del e
        torch.cuda.empty_cache()
    except Exception as ex:  # the extra configurations must never take the bench line down
        res["error"] = repr(ex)
    return res


def run_reference(args):
    rank, local_rank, world = dist_env()
    if rank != 0:
        return
    from atomorph_b200 import scenes
    from oracle import amref
    params = dict(seed=1, motion=amref.SPLINE, fading=amref.COSINE)
    images = scenes.square_to_disc(args.size)
    chains, blobs = reference_tables(args.size, images)
    m = build_reference(args.size, images, chains, blobs, params)
    cores = amref.hardware_concurrency()
    # swap leg (all host threads)
    m.set(threads=cores, cycle_length=100000)
    m.sync()
    sw = m.time_morph_steps(16)
    pps = 16 * max(1, cores) * 100000 / sw
    m.sync()
    # render leg: each step = ONE frame of the 64-frame morph (bounded sample; ~10-15 s per frame)
    warm = min(args.warmup, 1)
    for i in range(warm):
        m.time_render(0.5)
    t = 0.0
    for i in range(args.steps):
        dt, _ = m.time_render((i % args.frames) / float(args.frames))
        t += dt
    fps = args.steps / t
    line = {
        "impl": "reference", "metric": "morph frames/s at %d^2 RGBA (swap proposals/s in 'swap')" % args.size,
        "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
        "ms_per_step": 1000.0 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 synthetic %dx%d RGBA, 2 key frames, %d atoms, spline+cosine, %d frames" %
                   (args.size, args.size, chains[0]["words"].shape[1], args.frames), "step": "one frame (bounded sample)"},
        "swap": {"value": pps, "unit": "proposals/s", "cores": int(cores)},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 1, "kind": "reference",
                         "sample": "%d single frames of the %d-frame morph via am::morph::get_pixels (single-threaded renderer); "
                                   "swap: 16 morph steps x %d threads x 100000 proposals" % (args.steps, args.frames, cores)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    from atomorph_b200 import engine as eng
    from atomorph_b200 import scenes

    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(local_rank)
    peak, peak_kind = measured_peak()

    size, F = args.size, args.frames
    params = dict(seed=1, motion=eng.SPLINE, fading=eng.COSINE, threads=0, cycle_length=100000)
    images = scenes.square_to_disc(size)
    e = eng.Engine(local_rank, **params)          # private non-blocking stream: every kernel, copy AND collective of the engine runs on it
    e.load_images(images)
    e.step(8)                                     # blobify -> unify -> match -> init chains (all on device)
    assert e.state() == eng.STATE_ATOM_MORPHING
    info = e.chains() if size <= 256 else None
    A = e.table_device_ptr(0)[1]
    P = size * size
    # match first (untimed): the frames rendered below are those of the MATCHED morph.  cost ~ c_opt (1 + k/ppa) with
    # k ~ 110 on this scene: 24576 rounds = 12288 proposals per atom -> within 1 % of the optimum
    cost0 = e.cost()
    e.swap_rounds(args.match_rounds, want_stats=False)
    cost1 = e.cost()
    # north star: final transport cost no worse than 1 % above the reference's converged cost for the same seed and frames.
    # tests/golden/c2_copt.json = BASELINE.md section 3 protocol run on the unmodified reference (tests/golden/make_copt.py)
    copt = None
    try:
        with open(os.path.join(ROOT, "tests", "golden", "c2_copt.json")) as fh:
            g = json.load(fh)
        if int(g["size"]) == size:
            copt = float(g["c_opt"])
    except Exception:
        pass
    e.render_prepare()
    e.sync()

    # weak scaling: rank r renders frames [r*F, (r+1)*F) of an (F*world)-frame morph
    total_frames = F * world
    times = np.array([(rank * F + f) / float(total_frames) for f in range(F)])
    out = torch.empty((F, size, size), dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def render_step():
        e.render_into(times, out.data_ptr(), True)

    from atomorph_b200 import dist as amd
    if world > 1:
        amd.broadcast_table(e, rank, world, dev)                 # every rank renders / refines the same table (ncclBroadcast in the library)
        e.render_prepare()
    matcher = amd.ShardedMatcher(e, rank, world, device=dev, seed=1, p2p=os.environ.get("AMX_DIST_NCCL") is None)

    def swap_step():
        # weak scaling: per step every rank refines its 1/world of the atoms for `world` re-tiled epochs of 64 rounds
        # (= SWAP_ROUNDS/... proposals per rank as on one GPU), then ONE exchange of the column, issued by the library
        for _ in range(SWAP_EPOCHS):
            matcher.run_step(sub_epochs=world, rounds=64, column=1)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        tot = 0.0
        n0 = e.launch_count()
        for _ in range(steps):
            flush.zero_()                          # L2 flush between timed iterations (outside the event pair)
            torch.cuda.synchronize()
            e.timer_start()
            fn()
            tot += e.timer_stop()
        timed.launches += e.launch_count() - n0    # kernels of the engine launched inside the timed regions
        barrier()
        return tot
    timed.launches = 0

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_render = timed(render_step, args.steps, args.warmup)
    # per-kernel launch durations for the roofline: a second pass over the same steps with a CUDA event pair around every
    # render kernel launch (the pairs keep consecutive launches from overlapping their launch latencies, which costs ~8 %,
    # so `value` above is timed without them)
    counted = timed.launches
    e.kernel_times(True)
    timed(render_step, max(1, min(args.steps, 3)), 1)
    ktimes = e.kernel_times(False)                  # [k_bin2, k_acc]
    timed.launches = counted                        # (the diagnostic pass is not part of the timed region)
    # invariants of the sharded matcher, checked on every run at every N: each column stays the same multiset of key
    # points, the cost never rises (thread.cpp:1014-1038), and all replicas of the table are identical afterwards
    h = 2
    hash_before = [e.column_hash(j) for j in range(h)]
    cost_before = e.cost()
    st0 = e.swap_stats()
    ms_swap = timed(swap_step, args.steps, args.warmup)
    st1 = e.swap_stats()
    if world > 1:
        e.comm_check()
    cost_after = e.cost()
    hash_after = [e.column_hash(j) for j in range(h)]
    same_multiset = all(a[1] == b[1] for a, b in zip(hash_before, hash_after))
    cost_monotone = cost_after <= cost_before
    replicas_equal = True
    if world > 1:
        hv = torch.tensor([[v[0] >> 32, v[0] & 0xffffffff] for v in hash_after] + [[int(same_multiset), int(cost_monotone)]],
                          dtype=torch.int64, device=dev).reshape(-1)
        lo, hi = hv.clone(), hv.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        replicas_equal = bool(torch.equal(lo[:-2], hi[:-2]))
        same_multiset, cost_monotone = bool(lo[-2]), bool(lo[-1])
    e.render_prepare()                              # the e2e / parity legs below render the table the swap steps left
    clocks = sampler.stop()
    launches = timed.launches
    proposals = int(st1[0] - st0[0]) * args.steps // (args.steps + args.warmup)

    # e2e through the C-ABI with host buffers: upload the table (H2D), render, copy frames back (D2H)
    chain_words = e.chains()
    pinned_in = []
    for c in chain_words:                          # the step's inputs live in pinned host memory
        t = torch.from_numpy(np.ascontiguousarray(c["words"]).view(np.int64)).pin_memory()
        pinned_in.append(t)
        c["words"] = t.numpy().view(np.uint64)
    host_out = torch.empty((F, size, size), dtype=torch.int32).pin_memory()
    h2d = sum(c["words"].nbytes for c in chain_words)
    d2h = F * P * 4

    parts = {"h2d_table": 0.0, "render_prepare": 0.0, "render_and_d2h": 0.0}

    def e2e_step(record=False):
        t0 = time.perf_counter()
        e.import_chains(chain_words)                # H2D of the step's input (the trajectory table), synchronous
        t1 = time.perf_counter()
        e.render_prepare()                          # per-table sort + gather of the render inputs
        t2 = time.perf_counter()
        e.render_into(times, host_out.data_ptr(), False)   # render; frames travel D2H on a second stream while later batches render
        t3 = time.perf_counter()
        if record:
            parts["h2d_table"] += t1 - t0
            parts["render_prepare"] += t2 - t1
            parts["render_and_d2h"] += t3 - t2

    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        e2e_step(True)
    barrier()
    s_e2e = time.perf_counter() - t0
    e2e_parts = [1000.0 * parts[k] / e2e_steps for k in ("h2d_table", "render_prepare", "render_and_d2h")]
    # the drop-in call: am::morph::get_pixels(t, &vector) = amx_render_pixels, one frame of 8-byte am::pixel records per call
    # into pageable host memory (a std::vector in the reference's API), with and without the look-ahead ring
    facade = {}
    pix = np.zeros((size, size), dtype=np.uint64)
    for mode in ("lookahead", "one_frame_per_call"):
        e.set_lookahead(mode == "lookahead")
        for f in range(min(F, 16)):
            e.render_pixels_into(times[f], pix.ctypes.data)
        barrier()
        t0 = time.perf_counter()
        for f in range(F):
            e.render_pixels_into(times[f], pix.ctypes.data)
        facade[mode] = F / (time.perf_counter() - t0)
    e.set_lookahead(True)
    if world > 1:
        tf = torch.tensor([facade["lookahead"], facade["one_frame_per_call"]], dtype=torch.float64, device=dev)
        dist.all_reduce(tf, op=dist.ReduceOp.SUM)
        facade = {"lookahead": float(tf[0]), "one_frame_per_call": float(tf[1])}
    if world > 1:
        tp = torch.tensor(e2e_parts, dtype=torch.float64, device=dev)
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        e2e_parts = [float(v) for v in tp]

    # max over ranks
    if world > 1:
        t = torch.tensor([ms_render, ms_swap, s_e2e, float(proposals)], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_render, ms_swap, s_e2e = float(tmax[0]), float(tmax[1]), float(tmax[2])
        proposals = int(tsum[3])
    fps = world * F * args.steps / (ms_render / 1000.0)
    pps = proposals / (ms_swap / 1000.0)
    fps_e2e = world * F * e2e_steps / s_e2e

    render_bytes = A * 24 + P * 4                  # linear or h <= 3 -> 24 B/atom + 4 B/pixel
    swap_bytes = 32                                # h = 2: 4 distinct key points per proposal
    # roofline of the render batch = the kernel pair k_bin2 + k_acc (one launch each per batch of frames): algorithmic bytes
    # of the frames of a batch / the pair's device time, measured with CUDA events around each launch on the engine's stream
    kb, kt = ktimes
    if kb["launches"] and kt["launches"] and kt["frames"]:
        pair_ms = kb["ms"] / kb["launches"] + kt["ms"] / kt["launches"]
        frames_per_launch = kt["frames"] / float(kt["launches"])
        r_ach = render_bytes * frames_per_launch / (pair_ms / 1000.0) / 1e9
        kernels = {"k_bin2": {"us_per_launch": 1000.0 * kb["ms"] / kb["launches"], "launches": kb["launches"]},
                   "k_acc": {"us_per_launch": 1000.0 * kt["ms"] / kt["launches"], "launches": kt["launches"]},
                   "frames_per_launch": frames_per_launch}
    else:
        r_ach = render_bytes * (F * args.steps) / (ms_render / 1000.0) / 1e9
        kernels = None
    traffic = None
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r04_traffic.json")) as fh:
            traffic = json.load(fh).get("k_bin2+k_acc_dram_bytes_per_launch_pair")
    except Exception:
        pass
    s_ach = swap_bytes * (proposals / world) / (ms_swap / 1000.0) / 1e9
    # the tiled swap kernel runs out of shared memory and is bound by instruction issue, not by bytes: what it reaches of the
    # SM issue rate = warp-instructions per proposal (one ncu capture of the shipped kernel, profiles/r02_swap.json) x
    # proposals/s per GPU / (SMs x 4 schedulers x SM clock)
    swap_issue = None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_swap.json")) as fh:
            sw = json.load(fh)
        wipp = sw["warp_instructions_per_launch"] / float(sw["proposals_per_launch"])
        sm_hz = 1e6 * float(clocks.get("sm_mhz") or 1965.0)
        issue_peak = torch.cuda.get_device_properties(local_rank).multi_processor_count * 4 * sm_hz
        swap_issue = {"thread_instructions_per_proposal": 32.0 * wipp, "warp_instructions_per_proposal": wipp,
                      "issue_rate_frac": (proposals / world) / (ms_swap / 1000.0) * wipp / issue_peak,
                      "issue_peak_warp_instr_per_s": issue_peak, "ncu_issue_slots_busy_pct": sw.get("issue_slots_busy_pct"),
                      "source": "profiles/r02_swap.json, profiles/r02_ncu_swap.txt"}
    except Exception:
        pass

    line = {
        "metric": "morph frames/s at %d^2 RGBA (swap proposals/s in 'swap')" % size,
        "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_render / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 synthetic %dx%d RGBA, 2 key frames, %d atoms, spline+cosine, %d frames per GPU" % (size, size, A, F),
                   "l2": "256 MB buffer written between timed steps", "step": "render %d frames; swap: %d rounds" % (F, SWAP_ROUNDS),
                   "parallelism": "frames: frame-range x%d; swap: atom-range x%d + 1 all-gather/step" % (world, world)},
        "roofline": {"bound": "hbm", "achieved": r_ach, "peak": peak, "unit": "GB/s", "frac": r_ach / peak, "traffic": traffic,
                     "peak_kind": peak_kind, "kernel": "k_bin2+k_acc (one launch each per batch of frames)",
                     "bytes_per_unit": render_bytes, "unit_name": "frame", "kernels": kernels},
        "swap": {"value": pps, "unit": "proposals/s", "ms_per_step": ms_swap / args.steps, "rounds_per_step": SWAP_ROUNDS,
                 "issue": swap_issue,
                 "roofline": {"bound": "hbm", "achieved": s_ach, "peak": peak, "unit": "GB/s", "frac": s_ach / peak,
                              "traffic": None, "kernel": "k_swap_tiled", "bytes_per_unit": swap_bytes, "unit_name": "proposal",
                              "note": "NOT a bound for this kernel: tiles of 1024 atoms are refined for 64 rounds in shared memory per load, so the "
                                      "algorithmic 32 B/proposal never reach DRAM (~24 B per atom per 64 rounds do); see 'issue' for what limits it"}},
        "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step_max_over_ranks": {"h2d_table": e2e_parts[0], "render_prepare": e2e_parts[1], "render_and_d2h": e2e_parts[2]},
                "d2h_gbs_per_gpu": d2h / (e2e_parts[2] / 1000.0) / 1e9 if e2e_parts[2] > 0 else None, "host_numa": numa},
        "e2e_facade": {"value": facade["lookahead"], "unit": "frames/s", "call": "amx_render_pixels = am::morph::get_pixels(t, &vector): 8 B am::pixel records, one frame per call",
                       "one_frame_per_call": facade["one_frame_per_call"], "d2h_bytes_per_frame": int(P * 8)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "render_stats": dict(e.render_stats(), path_frames=e.render_path_frames(), tiled=e.render_tiled_stats()),
        "match": {"rounds": int(args.match_rounds), "proposals_per_atom": args.match_rounds / 2.0, "cost_initial": cost0, "cost_matched": cost1,
                  "reference_c_opt": copt, "cost_matched_over_c_opt": (cost1 / copt) if copt else None, "within_1pct_of_reference": (cost1 <= 1.01 * copt) if copt else None,
                  "cost_after_swap_steps": cost_after, "cost_monotone": cost_monotone, "columns_same_multiset": same_multiset, "replicas_equal": replicas_equal,
                  "exchange": ("p2p write-through + flag barrier" if matcher.p2p else "pack + ncclAllGather + unpack") if world > 1 else "none"},
    }

    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            from oracle import amref
            if amref.available():
                blobs = []
                for i in range(len(images)):
                    ys, xs = np.nonzero(images[i][..., 3] != 0)
                    labels, stats, meta = e.export_blobs(i)
                    blobs.append([dict(group=int(meta[0, 0]), stats=stats[0],
                                       surface=(ys.astype(np.uint64) * np.uint64(65536) + xs.astype(np.uint64)))])
                m = build_reference(size, images, chain_words, blobs, dict(seed=1, motion=eng.SPLINE, fading=eng.COSINE))
                cpu = cpu_reference_numbers(m, F, gpu_render=lambda t: e.render([t])[0])
                # full-size parity: the reference's frames against the GPU's at the same t on the same table (north star: bit-exact or <= 1 LSB)
                line["parity"] = dict(cpu["parity"], against="am::morph::get_pixels of oracle/_ref on the same chain table, %dx%d" % (size, size))
                line["cpu_baseline"] = {"value": cpu["fps"], "unit": "frames/s", "cores": 1, "kind": "reference",
                                        "sample": "%d of %d frames via am::morph::get_pixels on the same chain table (single-threaded renderer, %.1f s)"
                                                  % (cpu["frames"], F, cpu["render_s"]),
                                        "swap": {"value": cpu["pps"], "unit": "proposals/s", "cores": cpu["cores"],
                                                 "sample": "16 thread::morph steps x %d threads x 100000 proposals (%.1f s)" % (cpu["cores"], cpu["morph_s"])}}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
        except Exception as ex:  # the baseline must never take the GPU line down
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (ex,)}

    if rank == 0 and world == 1 and not args.no_configs:
        # the other BASELINE configurations, device-resident (N = 1): configs[2], [3], [4]
        del out, flush
        torch.cuda.empty_cache()
        line["configs"] = other_configs(dev, peak)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if not (same_multiset and cost_monotone and replicas_equal):
        raise SystemExit("sharded matcher invariant violated: %r" % (line["match"],))


# ------------------------------------------------------------------------------------------ B200 arm, BASELINE configs[4]
def run_b200_c5(args):
    """4096x4096 RGBA, 8 cyclic key frames (16.7 M atoms, 1 GiB trajectory table), 512 output frames.  STRONG scaling:
    matching is sharded by key-frame column phases (rank g refines every world-th column of the even / odd phase, then
    the refined columns go to every replica), rendering by contiguous output-frame range.  One step = one sweep of the
    matcher over all 8 columns (64 rounds per column) + the rank's share of the 512 frames."""
    import torch
    from atomorph_b200 import engine as eng
    from atomorph_b200 import scenes
    from atomorph_b200 import dist as amd

    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    peak, peak_kind = measured_peak()
    size = args.size if args.size != 1024 else 4096
    F, H = 512, 8
    images = scenes.rotating_shapes(size, H)
    e = eng.Engine(local_rank, seed=1, motion=eng.SPLINE, fading=eng.COSINE, threads=0, cycle_length=100000)
    e.load_images(images)
    del images
    e.step(8)
    assert e.state() == eng.STATE_ATOM_MORPHING
    A = e.table_device_ptr(0)[1]
    P = size * size
    matcher = amd.ShardedMatcher(e, rank, world, device=dev, seed=1, p2p=False)
    if world > 1:
        e.table_broadcast(0)
        matcher.p2p = os.environ.get("AMX_DIST_NCCL") is None and e.comm_enable_p2p()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cost0 = e.cost()
    for _ in range(4):                                  # untimed: a partly matched table (the tiles the renderer sees depend on it)
        matcher.run_sweep(epochs=2, rounds=64)
    cost1 = e.cost()
    hash_before = [e.column_hash(j) for j in range(H)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        matcher.run_sweep(epochs=1, rounds=64)
    st1 = e.swap_stats()
    barrier()
    ms_swap = 0.0
    n0 = e.launch_count()
    for _ in range(args.steps):
        e.timer_start()
        matcher.run_sweep(epochs=1, rounds=64)
        ms_swap += e.timer_stop()
        barrier()
    st2 = e.swap_stats()
    if world > 1:
        e.comm_check()
    cost2 = e.cost()
    hash_after = [e.column_hash(j) for j in range(H)]
    same_multiset = all(a[1] == b[1] for a, b in zip(hash_before, hash_after))
    cost_monotone = cost2 <= cost1
    replicas_equal = True
    if world > 1:
        hv = torch.tensor([[v[0] >> 32, v[0] & 0xffffffff] for v in hash_after] + [[int(same_multiset), int(cost_monotone)]],
                          dtype=torch.int64, device=dev).reshape(-1)
        lo, hi = hv.clone(), hv.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        replicas_equal = bool(torch.equal(lo[:-2], hi[:-2]))
        same_multiset, cost_monotone = bool(lo[-2]), bool(lo[-1])
    proposals = int(st2[0] - st1[0])                    # this rank's proposals in the timed sweeps

    # render: this rank's frame range, in chunks of 16 frames through one device buffer (inputs resident, far larger than L2)
    e.render_prepare()
    a, b = amd.frame_range(F, rank, world)
    CH = 16
    out = torch.empty((CH, size, size), dtype=torch.int32, device=dev)
    chunks = [np.arange(c, min(c + CH, b)) / float(F) for c in range(a, b, CH)]

    def render_pass():
        for ts in chunks:
            e.render_into(ts, out.data_ptr(), True)
    for _ in range(min(args.warmup, 1)):
        render_pass()
    barrier()
    ms_render = 0.0
    for _ in range(args.steps):
        e.timer_start()
        render_pass()
        ms_render += e.timer_stop()
        barrier()
    launches = e.launch_count() - n0
    # e2e: the same frames through the C-ABI into pinned HOST memory (one ring of 16 frames), table uploaded from the host first
    chain_words = e.chains()
    for c in chain_words:
        t = torch.from_numpy(np.ascontiguousarray(c["words"]).view(np.int64)).pin_memory()
        c["_pin"] = t
        c["words"] = t.numpy().view(np.uint64)
    host_out = torch.empty((CH, size, size), dtype=torch.int32).pin_memory()
    h2d = sum(c["words"].nbytes for c in chain_words)
    barrier()
    t0 = time.perf_counter()
    e.import_chains(chain_words)
    e.render_prepare()
    for ts in chunks:
        e.render_into(ts, host_out.data_ptr(), False)
    barrier()
    s_e2e = time.perf_counter() - t0
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms_render, ms_swap, s_e2e, float(proposals)], dtype=torch.float64, device=dev)
        tmax, tsum = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_render, ms_swap, s_e2e = float(tmax[0]), float(tmax[1]), float(tmax[2])
        proposals = int(tsum[3])
    fps = F * args.steps / (ms_render / 1000.0)
    pps = proposals / (ms_swap / 1000.0)
    render_bytes = A * 40 + P * 4                     # spline with h >= 4 -> 40 B/atom + 4 B/pixel
    r_ach = render_bytes * (b - a) * args.steps / (ms_render / 1000.0) / 1e9      # per GPU
    line = {
        "metric": "morph frames/s at %d^2 RGBA, %d cyclic key frames (swap proposals/s in 'swap')" % (size, H),
        "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_render / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C5 synthetic %dx%d RGBA, %d cyclic key frames, %d atoms, spline+cosine, %d frames in total" % (size, size, H, A, F),
                   "l2": "inputs (1 GiB table, 4 GiB sorted key points) and the output ring are larger than L2",
                   "step": "render %d frames per GPU; swap: one sweep = 64 rounds on each of the %d columns" % (b - a, H),
                   "parallelism": "frames: frame-range x%d; swap: key-frame columns of a phase x%d + exchange of the refined columns" % (world, world)},
        "roofline": {"bound": "hbm", "achieved": r_ach, "peak": peak, "unit": "GB/s", "frac": r_ach / peak, "traffic": None, "peak_kind": peak_kind,
                     "kernel": "k_bin2+k_acc (one launch each per batch of frames)", "bytes_per_unit": render_bytes, "unit_name": "frame", "per": "GPU"},
        "swap": {"value": pps, "unit": "proposals/s", "ms_per_step": ms_swap / args.steps,
                 "exchange": ("p2p write-through + flag barrier" if matcher.p2p else "ncclBroadcast per column") if world > 1 else "none"},
        "e2e": {"value": F / s_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int((b - a) * P * 4)},
        "gpu_launches": int(launches), "clocks": clocks,
        "render_stats": dict(e.render_stats(), path_frames=e.render_path_frames(), tiled=e.render_tiled_stats()),
        "match": {"cost_initial": cost0, "cost_after_8_sweeps": cost1, "cost_after_timed_sweeps": cost2, "cost_monotone": cost_monotone,
                  "columns_same_multiset": same_multiset, "replicas_equal": replicas_equal},
    }
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if not (same_multiset and cost_monotone and replicas_equal):
        raise SystemExit("sharded matcher invariant violated: %r" % (line["match"],))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of BASELINE configs 3, 4 and 5 (N = 1)")
    ap.add_argument("--match-rounds", type=int, default=49152, help="untimed pair-swap rounds before rendering (C2: 24576 proposals per atom)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"], help="c2: BASELINE configs[1] (the bench line); c5: configs[4], 4096^2 x 8 key frames, strong scaling")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_b200_c5(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
