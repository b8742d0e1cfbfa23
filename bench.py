#!/usr/bin/env python
"""bench.py — frames/s of the PillarNet point->BEV hot path (pillarize -> PFN -> sparse backbone ->
dense neck/head -> CenterHead decode -> NMS) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W              (N>1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the CPU arm: oracle port on host cores)

A step = one pass of the hot path over one batch of synthetic frames.  Workload at N=1: BASELINE.json
configs[1], PillarNet-18 nuScenes inference, batch 1 (`--workload nusc18`).  Frames are independent, so
N>1 shards frames across ranks with no data-path collective (weak scaling, fixed per-GPU work) and one
fixed-shape NCCL all-gather of the detections at the end of the job.

Prints ONE JSON line on rank 0 (see the keys in main()).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def _conv_bytes(r):
    """algorithmic bytes of one conv launch: bf16 activations in and out once, weights once"""
    if r.get("bytes") is not None:          # grouped head convs: n_groups planar 64-channel maps in, f32 columns out
        return int(r["bytes"])
    return int(2 * r["rows"] * (r["cin"] + r["cout"]) + 2 * r["taps"] * r["cin"] * r["cout"])


def _ncu_capture():
    for name in ("r2_ncu_full_convs_nusc18.json", "r1_ncu_full_convs_nusc18.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            return path, name
    return None, None


def _ncu_family_traffic(kernel):
    """average dram__bytes_read.sum + dram__bytes_write.sum PER LAUNCH over every launch of one kernel family in the
    committed `ncu --set full` capture of this command (one profiled pass, cold L2 per launch)"""
    path, name = _ncu_capture()
    if path is None:
        return None, None
    try:
        rows = [r for r in json.load(open(path)) if r["kernel"].startswith(kernel)
                and r.get("dram_read_MB") is not None]
        if not rows:
            return None, None
        tot = sum(r["dram_read_MB"] + r["dram_write_MB"] for r in rows) * 1e6
        return int(tot / len(rows)), f"profiles/{name}: mean of {len(rows)} {kernel} launches"
    except Exception:
        return None, None


def _ncu_traffic(top):
    """dram__bytes_read.sum + dram__bytes_write.sum of one conv shape from the committed `ncu --set full`
    capture of this same command (one profiled pass, cold L2 per launch):
    the launch of the same kernel template whose duration is closest."""
    path, cap_name = _ncu_capture()
    if path is None:
        return None, None
    try:
        rows = [r for r in json.load(open(path)) if r["kernel"].startswith(top["kernel"])]
        if top["kernel"] == "k_conv_dense":
            # the capture lists every template instance; pick by accumulated time of the family member that leads
            fam = {}
            for r in rows:
                fam.setdefault(r["kernel"], []).append(r)
            rows = max(fam.values(), key=lambda v: sum(x["us"] for x in v))
        if not rows:
            return None, None
        best = min(rows, key=lambda r: abs(r["us"] - 1.15 * top["avg_us"]))
        return int((best["dram_read_MB"] + best["dram_write_MB"]) * 1e6), f"profiles/{cap_name}: {best['kernel']}"
    except Exception:
        return None, None


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_model(workload, device, seed=0):
    import pillarnet_lts_b200 as P
    from pillarnet_lts_b200 import configs
    from pillarnet_lts_b200.registry import ConfigDict
    cfg = configs.get(workload)
    torch.manual_seed(seed)
    model = P.build_detector(ConfigDict.wrap(cfg["model"]), None, ConfigDict.wrap(cfg["test_cfg"]))
    g = torch.Generator().manual_seed(seed + 1)
    for m in model.modules():  # non-trivial BN statistics so the folded affine is exercised
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
    return model.to(device).eval(), cfg


def make_frames(kind, n, seed0):
    from pillarnet_lts_b200 import synth
    return synth.make_batch(kind, n, seed0)


# ---------------------------------------------------------------------------------------------------
def conv_breakdown(engine, reps=3):
    """Instrumented eager pass: CUDA events around every conv launch and every stage on the launching
    stream -> per-shape totals; used to name the dominant kernel and measure its duration live."""
    from pillarnet_lts_b200 import ops
    recs = []
    orig = ops.conv_gather

    inner = [1]  # identical back-to-back launches per event pair (conv pass: 4, amortises the cost of an event record)

    def timed(inp, weight, nbr, taps, cin, cout, out, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(inner[0]):
            r = orig(inp, weight, nbr, taps, cin, cout, out, **kw)
        e.record()
        # window-staged kernel (conv_win_tc.cu) when the caller passed tile plans for a raster-sorted submanifold
        # rulebook and the layer fits it (cout <= 256); gather kernel (conv_tcgen05.cu) otherwise
        kname = "k_conv_win" if (kw.get("nbr_plan") is not None and kw.get("nbr_kind") and cout <= 256
                                 and inp.dtype == torch.bfloat16 and out.dtype == torch.bfloat16) else "k_conv_tc"
        recs.append((taps, cin, cout, kw.get("rows_cap") or out.shape[0], kw.get("num"), s, e, kname, nbr, None))
        return r

    model = engine.model
    stages = {}

    def stage(name, fn):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = fn()
        e.record()
        stages.setdefault(name, []).append((s, e))
        return r

    orig_dense, orig_small = ops.conv_dense3x3, ops.conv3x3_small_cout

    def timed_dense(inp, in_coff, cin, n_frames, H, W, weight, cout, out, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(inner[0]):
            r = orig_dense(inp, in_coff, cin, n_frames, H, W, weight, cout, out, **kw)
        e.record()
        # FLOPs counted over real pixels only (the padded border rows are overhead)
        recs.append((9, cin, cout, n_frames * H * W, None, s, e, "k_conv_dense", None, None))
        return r

    def timed_small(inp, in_ld, cin, n_frames, H, W, groups, n_groups, wbuf, out, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(inner[0]):
            r = orig_small(inp, in_ld, cin, n_frames, H, W, groups, n_groups, wbuf, out, **kw)
        e.record()
        recs.append((9, cin, int(out.shape[1]), n_frames * H * W, None, s, e, "k_conv3x3_small", None, None))
        return r

    orig_grouped = ops.conv_dense3x3_grouped

    def timed_grouped(inp, in_coff, cin, n_groups, n_frames, H, W, weight, shift, group_tab, out, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(inner[0]):
            r = orig_grouped(inp, in_coff, cin, n_groups, n_frames, H, W, weight, shift, group_tab, out, **kw)
        e.record()
        # useful FLOPs: every group maps cin channels to its own few outputs (sum = packed output columns)
        recs.append((9, cin, int(out.shape[1]), n_frames * H * W, None, s, e, "k_conv_dense_grouped", None,
                     inp.numel() * 2 + out.numel() * 4 + weight.numel() * 2))
        return r

    orig_shift = ops.conv_dense3x3_grouped_shift

    def timed_shift(inp, n_groups, n_frames, H, W, weight, shift, group_tab, out):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(inner[0]):
            r = orig_shift(inp, n_groups, n_frames, H, W, weight, shift, group_tab, out)
        e.record()
        # same useful FLOPs as the implicit-GEMM form it replaces; the kernel itself is bound by reading `inp` once
        recs.append((9, 64, int(out.shape[1]), n_frames * H * W, None, s, e, "k_conv_shift", None,
                     inp.numel() * 2 + out.numel() * 4 + weight.numel() * 2))
        return r

    ops.conv_gather = timed
    ops.conv_dense3x3, ops.conv3x3_small_cout = timed_dense, timed_small
    ops.conv_dense3x3_grouped = timed_grouped
    ops.conv_dense3x3_grouped_shift = timed_shift
    try:
        with torch.cuda.stream(engine.stream), torch.no_grad():
            for rep in range(reps + 1):
                # passes 0..reps-1: stage timings with single launches; last pass: per-conv timings with 4 launches
                # per event pair (its stage times are discarded)
                inner[0] = 4 if rep == reps else 1
                if rep == reps:
                    recs.clear()
                    stage_keep = {k: list(v) for k, v in stages.items()}
                recs_start = len(recs)
                # park the stream for ~25 ms so the host enqueues the whole pass ahead of the device: the event
                # pairs then bracket back-to-back kernels (true device durations, no Python launch gaps)
                torch.cuda._sleep(50_000_000)
                sp = stage("reader", lambda: model.reader(dict(points_batched=(engine.points, engine.offsets))))
                feats = stage("backbone", lambda: model.backbone(sp))
                bev = stage("neck", lambda: model.neck(feats))
                preds = stage("head", lambda: model.bbox_head(bev))
                stage("decode_nms", lambda: model.bbox_head.predict_raw(preds, model.test_cfg))
            engine.stream.synchronize()
            stages = stage_keep
    finally:
        ops.conv_gather = orig
        ops.conv_dense3x3, ops.conv3x3_small_cout = orig_dense, orig_small
        ops.conv_dense3x3_grouped = orig_grouped
        ops.conv_dense3x3_grouped_shift = orig_shift
    per_pass = recs[recs_start:]
    shapes = {}
    pair_cache = {}
    for taps, cin, cout, rows_cap, num, s, e, kname, nbr, abytes in recs:
        rows = rows_cap if num is None else min(int(num.item()), rows_cap)
        # rulebook pairs P = present neighbours (SURVEY §8d counts a sparse conv as 2*P*Cin*Cout, not rows*taps: the
        # zero-filled taps of absent neighbours are not work)
        if nbr is not None:
            ck = (nbr.data_ptr(), rows)
            if ck not in pair_cache:
                pair_cache[ck] = int((nbr[:rows] >= 0).sum().item())
            pairs = pair_cache[ck]
        else:
            pairs = rows * taps
        key = (taps, cin, cout, rows, kname, pairs)
        d = shapes.setdefault(key, dict(us=0.0, n=0, bytes=abytes))
        d["us"] += s.elapsed_time(e) * 1e3 / 4
        d["n"] += 1
    out = []
    for (taps, cin, cout, rows, kname, pairs), d in shapes.items():
        flop = 2.0 * pairs * cin * cout
        avg = d["us"] / d["n"]
        out.append(dict(kernel=kname, taps=taps, cin=cin, cout=cout, rows=rows, pairs=pairs, bytes=d["bytes"],
                        launches_per_pass=d["n"], avg_us=avg, flop=flop,
                        tflops=flop / avg / 1e6, tflops_zero_filled=2.0 * rows * taps * cin * cout / avg / 1e6,
                        total_us_per_pass=d["us"]))
    out.sort(key=lambda r: -r["total_us_per_pass"])
    st = {k: float(np.median([s.elapsed_time(e) * 1e3 for s, e in v])) for k, v in stages.items()}
    return out, st, len(per_pass)


def parity_block(model, frames, B, pool, dev, precision):
    """Outside the timed region: the same frames through the fp32 mode (k_conv_simt, the 1e-3 parity path of
    tests/test_gpu_fullsize.py) and through the benchmarked mode; reports the worst head-map error relative to
    max|fp32| and how many detections of either run the other re-finds (same class, BEV IoU >= 0.7)."""
    import pillarnet_lts_b200 as P
    from pillarnet_lts_b200 import agreement
    dets, maps, plans = {}, {}, {}
    try:
        for prec in ("fp32", precision):
            P.set_precision(prec)
            dets[prec], maps[prec], plans[prec] = [], [], []
            with torch.no_grad():
                for i in range(pool):
                    fs = [torch.from_numpy(f).to(dev) for f in frames[i * B:(i + 1) * B]]
                    bev, _ = model.extract_feat(dict(points=fs))
                    preds = model.bbox_head(bev)
                    if i == 0:
                        maps[prec] = [{k: v.float().clone() for k, v in p.items()} for p in preds]
                    det_out, keep_count, plan = model.bbox_head.predict_raw(preds, model.test_cfg)
                    dets[prec] += model.bbox_head.assemble(det_out, keep_count, plan, None)
                    plans[prec].append(plan)
        torch.cuda.synchronize()
    finally:
        P.set_precision(precision)
    worst, worst_name = 0.0, None
    for t, (a, b) in enumerate(zip(maps[precision], maps["fp32"])):
        for k in b:
            e = agreement.rel_to_max(a[k], b[k])
            if e > worst:
                worst, worst_name = e, f"task{t}.{k}"
    cand = None
    for pa, pb in zip(plans["fp32"], plans[precision]):
        c = agreement.candidate_agreement(model.bbox_head, pa, pb, score_thr=float(model.test_cfg["score_threshold"]))
        if cand is None:
            cand = c
        else:
            for k in ("n_a", "n_b", "a_in_b", "b_in_a", "unexplained"):
                cand[k] += c[k]
            cand["mean_iou"] = (cand["mean_iou"] + c["mean_iou"]) / 2   # running mean over equally sized frames
            cand["min_iou"] = min(cand["min_iou"], c["min_iou"])
            cand["max_score_delta"] = max(cand["max_score_delta"], c["max_score_delta"])
    cand["recall_a_in_b"] = cand["a_in_b"] / max(1, cand["n_a"])
    cand["recall_b_in_a"] = cand["b_in_a"] / max(1, cand["n_b"])
    cand["note"] = ("unexplained = candidates listed by one run only whose score is NOT within 1e-3 of the other list's cut "
                    "(score threshold / pre_max-th score); 0 means the two runs differ only in near-ties at the cut")
    return {"against": "fp32 mode of the same model on the same frames (itself within 1e-3 of the dense-equivalent torch "
                       "restatement: tests/test_gpu_fullsize.py)",
            "frames": pool * B, "head_maps_max_rel_err": worst, "worst_map": worst_name,
            "candidates_pre_nms": cand,
            "note": "candidates = score-sorted top-pre_max boxes handed to NMS, matched by heat-map pixel + class; "
                    "detections = boxes kept by greedy NMS, matched at BEV IoU >= 0.7 (order-sensitive: near-tied "
                    "scores of random-init heads reorder the sweep)",
            "detections": agreement.summarize(dets["fp32"], dets[precision], iou_thr=0.7),
            "detections_top100": agreement.summarize(dets["fp32"], dets[precision], iou_thr=0.7, top=100)}


def run_gpu(args):
    import torch.distributed as dist
    import pillarnet_lts_b200 as P
    from pillarnet_lts_b200 import _lib
    from pillarnet_lts_b200.engine import StreamingEngine, calibrate_heatmap_bias

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"     # keep NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=dev)
    P.set_precision(args.precision)
    lib = _lib.load()
    B = args.frames_per_step
    cores = None
    if world > 1:
        from pillarnet_lts_b200.engine import pin_process_to_gpu_cores
        cores = pin_process_to_gpu_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    model, cfg = build_model(args.workload, dev)
    pool = 8
    # global frame i -> rank i % world (the reference's DistributedSampler: datasets/loader/sampler.py:93)
    frames = [make_frames(cfg["synth"], 1, seed0=1000 + (j * world + rank))[0] for j in range(pool * B)]
    calibrate_heatmap_bias(model, frames[:B], target_cells=args.hm_cells)
    cap = int(max(sum(len(f) for f in frames[i * B:(i + 1) * B]) for i in range(pool)) * 1.05) + 1024
    # `in_flight` lanes (stream + CUDA graph + buffers each) take the steps round-robin; lane 0 alone serves the
    # one-step-at-a-time legs (per-kernel breakdown, classic flushed timing, profiling pass)
    in_flight = args.in_flight if args.in_flight > 0 else (3 if B <= 2 else 1)
    n_lanes = 1 if args.profile_pass else in_flight
    seng = StreamingEngine(model, B, cap, in_flight=n_lanes, device=dev)
    seng.prepare(frames[:B], warmup=2)
    eng = seng.lanes[0]
    l0 = lib.pn_launch_count()
    with torch.cuda.stream(eng.stream), torch.no_grad():
        eng._forward()                      # one eager pass = the kernels one graph replay launches
    eng.stream.synchronize()
    launches_per_pass = lib.pn_launch_count() - l0

    # device-resident copies of the frame pool (for the HBM-resident `value` leg)
    dev_batches = []
    for i in range(pool):
        fs = frames[i * B:(i + 1) * B]
        offs = np.cumsum([0] + [len(f) for f in fs]).astype(np.int32)
        dev_batches.append((torch.from_numpy(np.concatenate(fs)).to(dev), torch.from_numpy(offs).to(dev)))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_resident(i):
        p, o = dev_batches[i % pool]
        with torch.cuda.stream(eng.stream):
            eng.points[:p.shape[0]].copy_(p, non_blocking=True)
            eng.offsets.copy_(o, non_blocking=True)
        eng.launch()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    # ---- value: inputs resident in HBM, L2 flushed between timed steps ----
    ev = []
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        with torch.cuda.stream(eng.stream):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(eng.stream)
        step_resident(i)
        with torch.cuda.stream(eng.stream):
            e.record(eng.stream)
        ev.append((s, e))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    gpu_ms = sum(s.elapsed_time(e) for s, e in ev)
    # ---- value with `in_flight` steps in flight: the same K steps, lanes round-robin, one device-timed bracket --------
    # Steps overlap, so there is no flush kernel between them; instead the steps cycle through a pool of distinct
    # frames larger than the 126 MB L2 (an input is re-read only after the whole pool went through).
    flight_ms, big_pool = None, None
    if n_lanes > 1:
        bytes_per_batch = max(1, sum(p.numel() * 4 for p, _ in dev_batches) // pool)
        n_big = int(min(64, max(pool, -(-160 * (1 << 20) // bytes_per_batch))))
        big = list(dev_batches)
        for j in range(pool, n_big):
            fs = [make_frames(cfg["synth"], 1, seed0=5000 + (j * B + b) * world + rank)[0] for b in range(B)]
            if sum(len(f) for f in fs) > cap:
                fs = [f[:cap // B] for f in fs]
            offs = np.cumsum([0] + [len(f) for f in fs]).astype(np.int32)
            big.append((torch.from_numpy(np.concatenate(fs)).to(dev), torch.from_numpy(offs).to(dev)))
        big_pool = {"batches": len(big), "bytes": int(sum(p.numel() * 4 for p, _ in big))}
        for i in range(max(args.warmup, 3) * n_lanes):
            seng.launch_resident(i, *big[i % len(big)])
        barrier()
        main = torch.cuda.current_stream(dev)
        seng.fork(main)
        for i in range(args.steps):
            seng.launch_resident(i, *big[i % len(big)])
        e0, e1 = seng.join(main)
        barrier()
        flight_ms = e0.elapsed_time(e1)
    # ---- e2e: pinned host -> H2D -> graph -> D2H of the detections, every step ----
    # public throughput API (InferenceEngine.run_pipelined): every step packs its frames into pinned memory,
    # copies them to the device, replays the graph and reads the detections back; the next step's packing + H2D
    # overlap the current step's compute (double-buffered input slots)
    staged = []
    for i in range(pool):
        staged.append([torch.from_numpy(f).pin_memory() for f in frames[i * B:(i + 1) * B]])
    run_e2e = seng.run if n_lanes > 1 else eng.run_pipelined
    run_e2e([staged[i % pool] for i in range(3 * n_lanes)])
    barrier()
    n_det_box = [0]

    def consume(i, dets):
        n_det_box[0] = int(sum(d["scores"].shape[0] for d in dets))

    t0 = time.perf_counter()
    _, h2d, d2h = run_e2e([staged[i % pool] for i in range(args.steps)], consume=consume)
    barrier()
    e2e_s = time.perf_counter() - t0
    # single-shot latency of the same public call, no overlap (reported next to the throughput figure)
    t0 = time.perf_counter()
    for i in range(min(args.steps, 10)):
        eng.infer(staged[i % pool])
    barrier()
    e2e_latency_ms = (time.perf_counter() - t0) * 1e3 / min(args.steps, 10)
    clocks = sampler.stop()
    # ---- value_sustained: >= args.sustain_s seconds of the same steps back to back (power / clock steady state) ----
    sustained = None
    if args.sustain_s > 0:
        est_ms = gpu_ms / args.steps + 0.06                     # step + L2 flush
        n_sus = max(args.steps, int(args.sustain_s * 1e3 / est_ms))
        sampler2 = ClockSampler(local)
        sampler2.start()
        ev2 = []
        for i in range(n_sus):
            with torch.cuda.stream(eng.stream):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(eng.stream)
            step_resident(i)
            with torch.cuda.stream(eng.stream):
                e.record(eng.stream)
            ev2.append((s, e))
            if (i & 255) == 255:
                ev2[-1][1].synchronize()                        # bound the queue depth
        barrier()
        sus_ms = sum(s.elapsed_time(e) for s, e in ev2)
        half = [s.elapsed_time(e) for s, e in ev2[n_sus // 2:]]
        sustained = {"steps": n_sus, "gpu_ms": sus_ms, "ms_per_step_second_half": float(np.mean(half)),
                     "clocks": sampler2.stop()}
        if n_lanes > 1:
            # the headline configuration (steps in flight, frame pool larger than L2) for the same duration
            n_sus = max(args.steps, int(args.sustain_s * 1e3 / (flight_ms / args.steps)))
            sampler3 = ClockSampler(local)
            sampler3.start()
            main = torch.cuda.current_stream(dev)
            seng.fork(main)
            for i in range(n_sus):
                lane = seng.launch_resident(i, *big[i % len(big)])
                if (i & 255) == 255:
                    lane.stream.synchronize()                   # bound the queue depth
            e0, e1 = seng.join(main)
            barrier()
            sustained.update({"in_flight_steps": n_sus, "in_flight_gpu_ms": e0.elapsed_time(e1),
                              "in_flight_clocks": sampler3.stop()})
    n_det = n_det_box[0]
    # ---- final detection gather (the only collective on the inference path) ----
    if world > 1:
        from pillarnet_lts_b200.dist import gather_detections
        gathered = gather_detections(eng.det_out, eng.keep_count)
        torch.cuda.synchronize()
    # max over ranks
    t = torch.tensor([gpu_ms, e2e_s * 1e3, sustained["gpu_ms"] if sustained else 0.0, flight_ms or 0.0,
                      sustained.get("in_flight_gpu_ms", 0.0) if sustained else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gpu_ms, e2e_ms, sus_ms, flight_ms, sus_flight_ms = t.tolist()
    frames_total = args.steps * B * world
    one_value = frames_total / (gpu_ms / 1e3)               # one step at a time, L2 flushed between steps
    one_ms_per_step = gpu_ms / args.steps
    value = frames_total / (flight_ms / 1e3) if n_lanes > 1 else one_value
    e2e_value = frames_total / (e2e_ms / 1e3)
    value_sustained = (sustained["steps"] * B * world / (sus_ms / 1e3)) if sustained else None
    if sustained and n_lanes > 1:
        sustained["one_in_flight_value"] = value_sustained
        value_sustained = sustained["in_flight_steps"] * B * world / (sus_flight_ms / 1e3)

    if args.profile_pass:
        # `ncu --profile-from-start off`: exactly one graph replay (one step) inside the profiler range
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident(0)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    # ---- extra: the other two B200 configurations of BASELINE.json, measured by every driver run -------------------
    # (config 3: PillarNet-34 Waymo, 8 frames per step STRONG-scaled over the ranks; config 4: PillarNet-34 nuScenes
    # training step, 4 frames per GPU, NCCL gradient all-reduce).  Outside the headline's timed regions.
    extra = None
    if not args.no_extra and not args.profile_pass:
        extra = {}
        try:
            per = max(1, 8 // world)
            r3 = measure_infer_resident("waymo34", per, 12, 3, args.precision, world, rank, dev)
            extra["config3_waymo34_8_frames_strong_scaled"] = {
                "value": r3["value"], "unit": "frames/s", "ms_per_step": r3["ms_per_step"], "frames_per_step_total": per * world,
                "frames_per_step_per_gpu": per, "scaling": "strong" if world <= 8 else "weak",
                "what": "device-resident frames/s, CUDA-graph replay, L2 flushed between steps, max over ranks"}
        except Exception as ex:                                   # the headline must survive a failure here
            extra["config3_waymo34_8_frames_strong_scaled"] = {"error": repr(ex)[:300]}
        try:
            r4 = measure_train("nusc34", 4, 8, 3, args.precision, world, rank, dev)
            extra["config4_nusc34_training_4_frames_per_gpu"] = {
                "value": r4["value"], "unit": "frames/s", "ms_per_step": r4["ms_per_step"], "launches_per_step": r4["launches"] // 8,
                "scaling": "weak", "loss_last": r4["loss_last"],
                "what": "TrainEngine step (forward+loss+backward graph, gradient all-reduce, optimiser graph), "
                        "max over ranks"}
        except Exception as ex:
            extra["config4_nusc34_training_4_frames_per_gpu"] = {"error": repr(ex)[:300]}
        P.set_precision(args.precision)
    line = None
    if rank == 0:
        peaks = _peaks()
        breakdown, stages, n_conv = conv_breakdown(eng)
        step_us = gpu_ms * 1e3 / args.steps
        # Peak the conv kernels are held against: they are timed one shape at a time (4 back-to-back launches between
        # two events of an eager pass, warm L2), i.e. in isolation, so the BURST bf16 figure applies whenever the SM
        # clock sampled during the run sat at its maximum; the sustained figure otherwise.
        at_max = bool(clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.97 * clocks["sm_max_mhz"])
        tf_peak = peaks["tf_burst"] if at_max else peaks["tf_sustained"]
        tf_peak_name = peaks["source"] + (" (burst bf16: kernels timed in isolation at max SM clock)" if at_max
                                          else " (sustained bf16: SM clock below max during the run)")
        fams = {}
        fam_of = {"k_conv_win": "sparse convs (k_conv_win + k_conv_tc)", "k_conv_tc": "sparse convs (k_conv_win + k_conv_tc)"}
        for r in breakdown:
            f = fams.setdefault(fam_of.get(r["kernel"], r["kernel"]), dict(us=0.0, flop=0.0, launches=0, bytes=0))
            f["us"] += r["total_us_per_pass"]
            f["flop"] += r["flop"] * r["launches_per_pass"]
            f["launches"] += r["launches_per_pass"]
            f["bytes"] += _conv_bytes(r) * r["launches_per_pass"]
        families = [dict(kernel=k, launches_per_pass=f["launches"], total_us_per_pass=f["us"],
                         share_of_step=f["us"] / step_us, tflops=f["flop"] / f["us"] / 1e6,
                         frac=f["flop"] / f["us"] / 1e6 / tf_peak,
                         algorithmic_gbs=f["bytes"] / f["us"] / 1e3,
                         algorithmic_bytes_per_launch=int(f["bytes"] / max(1, f["launches"])))
                    for k, f in fams.items()]
        families.sort(key=lambda r: -r["total_us_per_pass"])
        fam = families[0]
        top = breakdown[0]
        flop_total = sum(r["flop"] * r["launches_per_pass"] for r in breakdown)
        nusc_b1 = args.workload == "nusc18" and B == 1
        traffic, traffic_src = (_ncu_family_traffic("k_conv_win" if fam["kernel"].startswith("sparse") else fam["kernel"])
                                if nusc_b1 else (None, None))
        top_traffic, top_traffic_src = _ncu_traffic(top) if nusc_b1 else (None, None)
        # `roofline` = the kernel FAMILY with the largest share of the step (all its launches: sum of 2*P*Cin*Cout over
        # sum of durations); `best_shape` = the single conv shape with the largest share, for comparison
        roof = {"bound": "tensor", "kernel": fam["kernel"] + " (tcgen05 implicit-GEMM convs, all launches of the family; the "
                                                             "sparse family also holds the strided and the two dense gather "
                                                             "convs that share k_conv_tc)",
                "achieved": fam["tflops"], "peak": tf_peak, "unit": "TFLOP/s", "frac": fam["frac"],
                "peak_source": tf_peak_name, "launches_per_step": fam["launches_per_pass"],
                "avg_us": fam["total_us_per_pass"] / fam["launches_per_pass"], "share_of_step": fam["share_of_step"],
                "flops": "2*P*Cin*Cout with P = rulebook pairs actually present (dense convs: P = 9*pixels)",
                "share_basis": "share_of_step = kernel time / the one_in_flight step (one pass of one frame); with steps "
                               "in flight the conv families, which fill every SM, add up to the whole step",
                "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write, family mean)",
                "traffic_source": traffic_src, "algorithmic_bytes": fam["algorithmic_bytes_per_launch"],
                "families": families,
                "best_shape": {"kernel": top["kernel"], "shape": {k: top[k] for k in ("taps", "cin", "cout", "rows")},
                               "achieved": top["tflops"], "frac": top["tflops"] / tf_peak, "avg_us": top["avg_us"],
                               "share_of_step": top["total_us_per_pass"] / step_us, "traffic": top_traffic,
                               "traffic_source": top_traffic_src, "algorithmic_bytes": _conv_bytes(top)}}
        parity = parity_block(model, frames, B, pool, dev, args.precision) if not args.no_parity else None
        cpu = cpu_baseline(args, model, frames[0:B]) if world == 1 and not args.no_cpu_baseline else None
        line = {
            "metric": "frames/s (pillarize->PFN->sparse backbone->dense neck/head->decode->NMS)",
            "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": (flight_ms if n_lanes > 1 else gpu_ms) / args.steps,
            "ms_per_frame": (flight_ms if n_lanes > 1 else gpu_ms) / args.steps / B,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "bf16x3": "bf16x3 (three bf16 tensor-core products per fp32 product, fp32 "
                                                "accumulate and activations)"}.get(args.precision, "f32"),
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: PillarNet inference, batch {B}/GPU, synthetic "
                                   f"{cfg['synth']}-shaped frames (~{int(np.mean([len(f) for f in frames]))} pts), "
                                   f"random-init weights, hm bias calibrated to ~{args.hm_cells} candidate cells/task",
                       "frames_per_step_per_gpu": B,
                       "steps_in_flight_per_gpu": n_lanes,
                       "l2": (f"steps cycle through {big_pool['batches']} distinct resident batches = "
                              f"{big_pool['bytes'] / 1e6:.0f} MB, larger than the 126 MB L2 (steps overlap, so no flush "
                              "kernel between them); one_in_flight: 256 MB buffer written between timed steps"
                              if n_lanes > 1 else "256 MB buffer written between timed steps (L2 flush)"),
                       "timing": ("one CUDA-event bracket around the K steps (lanes fork from / join into the timing "
                                  "stream), max over ranks" if n_lanes > 1 else
                                  "CUDA events around every step on the launching stream, summed, max over ranks"),
                       "cuda_graph": True, "precision": args.precision},
            # the same K steps one at a time on one lane with an L2 flush between steps (the round-1 / early round-2
            # headline method): per-step latency and the base of the per-kernel shares in `roofline`
            "one_in_flight": {"value": one_value, "unit": "frames/s", "ms_per_step": one_ms_per_step,
                              "l2": "256 MB buffer written between timed steps (L2 flush)"},
            "clocks": clocks,
            "value_sustained": value_sustained,
            "sustained": ({"seconds": sus_ms / 1e3, "steps": sustained["steps"],
                           "ms_per_step": sus_ms / sustained["steps"],
                           "ms_per_step_second_half": sustained["ms_per_step_second_half"],
                           "clocks": sustained["clocks"],
                           "one_in_flight_value": sustained.get("one_in_flight_value"),
                           "in_flight_steps": sustained.get("in_flight_steps"),
                           "in_flight_seconds": sus_flight_ms / 1e3 if n_lanes > 1 else None,
                           "in_flight_clocks": sustained.get("in_flight_clocks"),
                           "note": "the timed steps repeated back to back for this long (one at a time with flushes: "
                                   "seconds / steps / ms_per_step; and in the headline configuration: in_flight_*): the "
                                   "clock / power steady state of a streaming deployment"}
                          if sustained else None),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "single_shot_latency_ms": e2e_latency_ms,
                    "api": ("StreamingEngine.run" if n_lanes > 1 else "InferenceEngine.run_pipelined") +
                           " (frames handed over in pinned host memory; per lane the next batch's H2D overlaps the replay)",
                    "host_cores_rank0": cores},
            "gpu_launches": int(launches_per_pass * args.steps),
            "launches_per_step": int(launches_per_pass),
            "wall_s_timed_region": t_wall,
            "detections_last_step": n_det,
            "roofline": roof,
            "parity": parity,
            "model_tflops": flop_total / step_us / 1e6,
            "stages_us": stages,
            "cpu_baseline": cpu,
            "extra": extra,
        }
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        tag = "_prof" if args.profile_pass else ""
        with open(os.path.join(ROOT, "gpurun_out", f"breakdown_{args.workload}_n{world}{tag}.json"), "w") as fh:
            json.dump({"convs": breakdown, "stages_us": stages, "conv_launches_per_pass": n_conv}, fh, indent=1)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


# ---------------------------------------------------------------------------------------------------
def measure_train(workload, B, steps, warmup, precision, world, rank, dev, eager=False):
    """one data-parallel training configuration (BASELINE config 4): returns a dict of results (every rank)"""
    import torch.distributed as dist
    import pillarnet_lts_b200 as P
    from pillarnet_lts_b200 import _lib, configs, train
    from pillarnet_lts_b200.dist import GradientAverager
    from pillarnet_lts_b200.registry import ConfigDict
    P.set_precision(precision)
    lib = _lib.load()
    cfg = configs.get(workload)
    torch.manual_seed(0)
    model = P.build_detector(ConfigDict.wrap(cfg["model"]), cfg["train_cfg"], ConfigDict.wrap(cfg["test_cfg"]))
    model = model.to(dev).train()
    # fused multi-tensor AdamW: one kernel per parameter group instead of a dozen foreach passes over 515 tensors
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01, capturable=not eager, fused=True)
    avg = GradientAverager(list(model.parameters()), bucket_mb=25, module=model) if world > 1 else None
    pool = 4
    rng = np.random.default_rng(7 + rank)
    H, W = model.reader.height, model.reader.width
    batches = []
    for i in range(pool):
        fs = [make_frames(cfg["synth"], 1, seed0=2000 + ((i * B + j) * world + rank))[0] for j in range(B)]
        offs = np.cumsum([0] + [len(f) for f in fs]).astype(np.int32)
        ex = {"points_batched": (torch.from_numpy(np.concatenate(fs)).to(dev), torch.from_numpy(offs).to(dev)),
              "points": None, "metadata": [None] * B}
        ex.update(train.synthetic_targets(model.bbox_head, B, H, W, rng, max_objs=500, device=dev))
        batches.append(ex)
    n_pts = int(np.mean([b["points_batched"][0].shape[0] for b in batches]) / B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = None
    if not eager:
        # sync-free step as two CUDA graphs (train.TrainEngine): fixed input buffers, live row counts on the device
        cap = int(max(b["points_batched"][0].shape[0] for b in batches) * 1.02) + 1024
        eng = train.TrainEngine(model, opt, B, cap, batches[0], averager=avg).prepare(warmup=3)
        step = lambda i: eng.step(batches[i % pool])
        stream = eng.stream
    else:
        step = lambda i: train.train_step(model, batches[i % pool], opt, avg)
        stream = torch.cuda.current_stream()
    for i in range(max(warmup, 3)):
        loss = step(i)
    barrier()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    l0 = lib.pn_launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream)
    for i in range(steps):
        loss = step(i)
    e.record(stream)
    barrier()
    gpu_ms = s.elapsed_time(e)
    launches = lib.pn_launch_count() - l0
    if eng is not None:
        launches = eng.launches_per_step * steps
    clocks = sampler.stop()
    t = torch.tensor([gpu_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gpu_ms = t.item()
    res = dict(value=steps * B * world / (gpu_ms / 1e3), ms_per_step=gpu_ms / steps, launches=int(launches),
               loss_last=float(loss), clocks=clocks, n_pts=n_pts, train_state=train.is_static())
    train.set_static(False)
    del eng, model, opt, batches
    torch.cuda.empty_cache()
    return res


def measure_infer_resident(workload, B, steps, warmup, precision, world, rank, dev):
    """device-resident frames/s of one inference configuration, frames sharded rank-major (i -> rank i % world); used
    for the strong-scaling leg of BASELINE config 3 (8 Waymo-shaped frames per step over all ranks)"""
    import torch.distributed as dist
    import pillarnet_lts_b200 as P
    from pillarnet_lts_b200.engine import InferenceEngine, calibrate_heatmap_bias
    P.set_precision(precision)
    model, cfg = build_model(workload, dev)
    pool = 4
    frames = [make_frames(cfg["synth"], 1, seed0=3000 + (j * world + rank))[0] for j in range(pool * B)]
    calibrate_heatmap_bias(model, frames[:B], target_cells=1500)
    cap = int(max(sum(len(f) for f in frames[i * B:(i + 1) * B]) for i in range(pool)) * 1.05) + 1024
    eng = InferenceEngine(model, B, cap, device=dev)
    eng.upload(eng.stage_host(frames[:B]))
    eng.prepare(warmup=2)
    dev_batches = []
    for i in range(pool):
        fs = frames[i * B:(i + 1) * B]
        offs = np.cumsum([0] + [len(f) for f in fs]).astype(np.int32)
        dev_batches.append((torch.from_numpy(np.concatenate(fs)).to(dev), torch.from_numpy(offs).to(dev)))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(i):
        p, o = dev_batches[i % pool]
        with torch.cuda.stream(eng.stream):
            eng.points[:p.shape[0]].copy_(p, non_blocking=True)
            eng.offsets.copy_(o, non_blocking=True)
        eng.launch()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(warmup, 3)):
        step(i)
    barrier()
    ev = []
    for i in range(steps):
        with torch.cuda.stream(eng.stream):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(eng.stream)
        step(i)
        with torch.cuda.stream(eng.stream):
            e.record(eng.stream)
        ev.append((s, e))
    barrier()
    gpu_ms = sum(s.elapsed_time(e) for s, e in ev)
    t = torch.tensor([gpu_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gpu_ms = t.item()
    del eng, model, dev_batches, flush
    torch.cuda.empty_cache()
    return dict(value=steps * B * world / (gpu_ms / 1e3), ms_per_step=gpu_ms / steps, frames_per_step_per_gpu=B)


def run_train(args):
    """BASELINE config 4: one data-parallel training step (reader + sparse backbone forward/backward on the
    library's kernels, dense neck/head + loss in PyTorch, NCCL gradient all-reduce)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    B = args.frames_per_step
    r = measure_train(args.workload, B, args.steps, args.warmup, args.precision, world, rank, dev, eager=args.train_eager)
    if rank == 0:
        line = {
            "metric": "frames/s (training step: reader + sparse backbone fwd/bwd, dense neck/head + loss, grad all-reduce)",
            "value": r["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}-train: {B} frames/GPU (~{r['n_pts']} pts/frame), AdamW, synthetic targets, "
                                   f"random-init weights; inputs larger than L2 (activations ~GBs per step)",
                       "frames_per_step_per_gpu": B, "precision": args.precision,
                       "mode": ("eager, exactly sized rows, one host sync per rulebook (as the reference)" if args.train_eager
                                else "TrainEngine: forward+loss+backward and the optimiser step as two CUDA graphs, no "
                                     "host sync; input batch copied device-to-device into fixed buffers inside the "
                                     "timed region")},
            "clocks": r["clocks"], "gpu_launches": r["launches"], "loss_last": r["loss_last"],
            "e2e": None, "roofline": None, "cpu_baseline": None,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
def cpu_run(model, frames, steps, warmup, pillarizer="numpy"):
    """the oracle port of the whole path on the host cores; returns (seconds per step list, timings)"""
    import copy
    from oracle import cpu_path
    model_cpu = copy.deepcopy(model).cpu().eval()
    times, tm = [], {}
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        cpu_path.cpu_forward(model_cpu, frames, tm, pillarizer=pillarizer)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, tm


def _cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return None


def cpu_config1(model, frame, reps=7):
    """BASELINE.json configs[0]: the reference's numba `points_to_voxel` (det3d/ops/point_cloud/point_cloud_ops.py:
    112-184, imported by path from oracle/_ref/py; one thread, JIT warm-up excluded) + the PFN forward on torch-CPU
    (Linear(7,32) + BN1d + ReLU + amax scatter) for ONE frame; median of `reps`."""
    from oracle import build_ref, cpu_path
    if not build_ref.python_available():
        return None
    import copy
    rd = copy.deepcopy(model.reader).cpu().eval().pfn_layers
    cpu_path.numba_reader(rd, [frame])                       # JIT compile + first-touch, untimed
    tv, tp = [], []
    for _ in range(reps):
        _, t = cpu_path.numba_reader(rd, [frame])
        tv.append(t["points_to_voxel_s"])
        tp.append(t["pfn_s"])
    sec = float(np.median(np.array(tv) + np.array(tp)))
    return {"value": 1.0 / sec, "unit": "frames/s", "kind": "reference",
            "what": "numba points_to_voxel (max_points 64, max_voxels 150000; 1 thread) + torch-CPU PFN forward, one "
                    f"frame of {len(frame)} points, median of {reps} after a JIT warm-up",
            "points_to_voxel_ms": float(np.median(tv)) * 1e3, "pfn_ms": float(np.median(tp)) * 1e3,
            "cores": {"points_to_voxel": 1, "pfn": torch.get_num_threads()}}


def cpu_baseline(args, model, frames):
    torch.set_num_threads(os.cpu_count() or 1)
    from oracle import build_ref
    pillarizer = "numba" if build_ref.python_available() else "numpy"
    times, tm = cpu_run(model, frames, steps=5, warmup=1, pillarizer=pillarizer)
    sec = float(np.median(times))
    return {"value": len(frames) / sec, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
            "cpu": _cpu_model_name(), "host_cores": os.cpu_count(),
            "sample": f"{len(frames)} full synthetic frame(s) per step, 1 warm-up + 5 timed steps (median) of the "
                      f"reference's CPU formulation of the path: "
                      + ("the reference's own numba points_to_voxel" if pillarizer == "numba" else "numpy pillarize")
                      + ", torch-CPU PFN + dense-equivalent backbone + neck/head, C NMS",
            "seconds_per_step": sec, "seconds_per_step_all": times, "stage_seconds": tm,
            "config1_pillarize_pfn": cpu_config1(model, frames[0])}


def run_reference(args):
    """--impl reference: the CPU arm. Rank 0 alone runs; other ranks exit 0 without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    B = args.frames_per_step
    model, cfg = build_model(args.workload, torch.device("cpu"))
    frames = make_frames(cfg["synth"], B, seed0=1000)
    # bounded: each step is B full frames; steps/warm-up are clamped so the run ends within minutes
    from oracle import build_ref
    pillarizer = "numba" if build_ref.python_available() else "numpy"
    t0 = time.perf_counter()
    times, tm = cpu_run(model, frames, steps=1, warmup=0, pillarizer=pillarizer)   # also the numba JIT warm-up
    est = times[0]
    budget = 150.0
    steps = max(1, min(args.steps, int(budget / max(est, 1e-3))))
    warm = max(0, min(args.warmup, 1))
    times, tm = cpu_run(model, frames, steps=steps, warmup=warm, pillarizer=pillarizer)
    total = float(np.sum(times))
    value = steps * B / total
    line = {
        "impl": "reference",
        "metric": "frames/s (pillarize->PFN->sparse backbone->dense neck/head->decode->NMS)",
        "value": value, "unit": "frames/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": steps, "warmup": warm, "steps_requested": args.steps,
        "ms_per_step": total / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: PillarNet inference, batch {B}, CPU oracle port of the reference "
                               f"path on the host cores (no GPU)", "frames_per_step_per_gpu": B},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                         "cpu": _cpu_model_name(), "host_cores": os.cpu_count(),
                         "sample": f"{B} full synthetic frame(s) per step; steps clamped to fit ~{budget:.0f} s; "
                                   f"pillarization: " + ("reference numba points_to_voxel" if pillarizer == "numba"
                                                         else "numpy port"),
                         "stage_seconds": tm, "seconds_per_step_all": times,
                         "config1_pillarize_pfn": cpu_config1(model, frames[0])},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def _protect_stdout():
    """stdout carries exactly ONE JSON line (the driver parses it).  Libraries write there too (NCCL prints its version
    banner at NCCL_DEBUG >= VERSION, which launchers set): from here on file descriptor 1 points at stderr and the
    result line goes to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="nusc18", choices=["nusc18", "nusc34", "waymo34"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="train: BASELINE config 4 (training step); not the headline metric")
    ap.add_argument("--frames-per-step", type=int, default=1)
    ap.add_argument("--in-flight", type=int, default=0,
                    help="steps in flight per GPU (StreamingEngine lanes); 1 = one step at a time with L2 flushes; "
                         "0 = auto: 3 for steps of one or two frames (whose tail kernels leave most SMs idle), 1 for "
                         "larger batches (measured: waymo34 batch 8 gains nothing)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "bf16x3"],
                    help="bf16: the fast mode (headline); bf16x3: split-bf16 tensor-core mode at fp32-grade accuracy; "
                         "fp32: FMA kernels (validation)")
    ap.add_argument("--hm-cells", type=int, default=1500)
    ap.add_argument("--train-eager", action="store_true",
                    help="--mode train: the eager, exactly sized path (host syncs) instead of the CUDA-graph TrainEngine")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the extra block (config 3 strong scaling, config 4 training step)")
    ap.add_argument("--no-parity", action="store_true", help="skip the fp32-vs-bf16 parity block (outside the timed region)")
    ap.add_argument("--sustain-s", type=float, default=2.5,
                    help="length of the back-to-back sustained leg in seconds (0 disables; reported as value_sustained)")
    ap.add_argument("--profile-pass", action="store_true",
                    help="wrap one extra step in cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; see --impl reference)")
    if args.mode == "train":
        run_train(args)
        return
    run_gpu(args)


if __name__ == "__main__":
    main()
