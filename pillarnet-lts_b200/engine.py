"""Inference engine: the whole point->detections pass as one CUDA graph with device-resident counts.

The reference spends its time on ~500 launches and ~20 host syncs per frame; here the
forward pass (pn_pillarize ... pn_nms) is sync-free, so it is captured once per (batch, capacity) and
replayed.  Public API (what bench.py's `e2e` leg times):

    eng = InferenceEngine(model, n_frames=1, points_cap=300_000)
    dets = eng.infer([points_f32_numpy_or_tensor, ...])      # H2D -> graph -> D2H -> list of dicts

`infer` accepts frames of any size up to the capacity: the live point count travels in the
frame-offset vector on the device.
"""
import numpy as np
import torch

from . import config


class InferenceEngine:
    def __init__(self, model, n_frames, points_cap, point_dim=5, device=None, use_graph=True):
        self.model = model.eval()
        self.B = n_frames
        self.cap = int(points_cap)
        self.dev = device or next(model.parameters()).device
        self.points = torch.zeros(self.cap, point_dim, dtype=torch.float32, device=self.dev)
        self.offsets = torch.zeros(n_frames + 1, dtype=torch.int32, device=self.dev)
        self.h_points = torch.zeros(self.cap, point_dim, dtype=torch.float32).pin_memory()
        self.h_offsets = torch.zeros(n_frames + 1, dtype=torch.int32).pin_memory()
        self.graph = None
        self.use_graph = use_graph
        self.det_out = self.keep_count = self.plan = None
        self.h_det = self.h_cnt = None
        self.stream = torch.cuda.Stream(device=self.dev)

    # -- input staging -------------------------------------------------------------------------
    def stage_host(self, frames):
        """packs frames into the pinned host buffers; returns the number of points"""
        assert len(frames) == self.B
        n = 0
        self.h_offsets[0] = 0
        for b, f in enumerate(frames):
            f = torch.as_tensor(f, dtype=torch.float32)
            k = f.shape[0]
            if n + k > self.cap:
                raise RuntimeError(f"{n + k} points exceed the engine capacity {self.cap}")
            self.h_points[n:n + k].copy_(f)
            n += k
            self.h_offsets[b + 1] = n
        return n

    def upload(self, n):
        """async H2D of the staged frames (pinned -> device) on the engine stream"""
        with torch.cuda.stream(self.stream):
            self.points[:n].copy_(self.h_points[:n], non_blocking=True)
            self.offsets.copy_(self.h_offsets, non_blocking=True)
        return n * self.points.shape[1] * 4 + self.offsets.numel() * 4

    # -- graph ---------------------------------------------------------------------------------
    def _forward(self):
        return self.model.forward_device(self.points, self.offsets)

    def prepare(self, warmup=2):
        """eager warm-up (fills lowering caches, static tables; the first pass also measures the active-row
        counts of the sparse stages on whatever batch is staged, which steer the conv tile shapes) then capture"""
        from . import backbone
        with torch.cuda.stream(self.stream), torch.no_grad():
            for i in range(max(1, warmup)):
                if i == 0:
                    backbone.observe_rows_begin()
                out = self._forward()
                if i == 0:
                    self.stream.synchronize()
                    self.rows_seen = backbone.observe_rows_end()
            self.stream.synchronize()
            if self.use_graph:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=self.stream):
                    out = self._forward()
            self.det_out, self.keep_count, self.plan = out
        self.h_det = torch.empty(self.det_out.shape, dtype=self.det_out.dtype).pin_memory()
        self.h_cnt = torch.empty(self.keep_count.shape, dtype=self.keep_count.dtype).pin_memory()
        self.stream.synchronize()
        return self

    def launch(self):
        """one forward pass on the engine stream (no host sync)"""
        with torch.cuda.stream(self.stream), torch.no_grad():
            if self.graph is not None:
                self.graph.replay()
            else:
                self.det_out, self.keep_count, self.plan = self._forward()

    def download(self):
        with torch.cuda.stream(self.stream):
            self.h_det.copy_(self.det_out, non_blocking=True)
            self.h_cnt.copy_(self.keep_count, non_blocking=True)
        return self.h_det.numel() * 4 + self.h_cnt.numel() * 4

    # -- public API ------------------------------------------------------------------------------
    def infer(self, frames, metadata=None):
        n = self.stage_host(frames)
        self.upload(n)
        if self.det_out is None:
            self.prepare()       # after the upload: the warm-up passes see a real batch
        self.launch()
        self.download()
        self.stream.synchronize()
        return self.assemble_host(metadata)

    # -- throughput API: double-buffered input, copies overlapped with compute ----------------------------------
    def _init_pipeline(self):
        """two input slots (pinned host + device staging) and two pinned result slots; H2D runs on its own
        stream, the graph's fixed input buffers are refreshed by a device-to-device copy (a few microseconds)"""
        D = self.points.shape[1]
        self._p_host = [(torch.zeros(self.cap, D).pin_memory(), torch.zeros(self.B + 1, dtype=torch.int32).pin_memory())
                        for _ in range(2)]
        self._p_dev = [(torch.zeros(self.cap, D, device=self.dev), torch.zeros(self.B + 1, dtype=torch.int32,
                                                                                device=self.dev)) for _ in range(2)]
        self._p_res = [(torch.empty(self.det_out.shape, dtype=self.det_out.dtype).pin_memory(),
                        torch.empty(self.keep_count.shape, dtype=self.keep_count.dtype).pin_memory()) for _ in range(2)]
        self._copy_stream = torch.cuda.Stream(device=self.dev)
        self._ev_up = [torch.cuda.Event() for _ in range(2)]
        self._ev_used = [torch.cuda.Event() for _ in range(2)]
        self._ev_done = [torch.cuda.Event() for _ in range(2)]
        self._n = [0, 0]

    def _prefetch(self, slot, frames):
        """host pack + async H2D of one batch into input slot `slot`; returns the bytes copied"""
        hp, ho = self._p_host[slot]
        dp, do = self._p_dev[slot]
        frames = [torch.as_tensor(f, dtype=torch.float32) for f in frames]
        total = sum(f.shape[0] for f in frames)
        if total > self.cap:
            raise RuntimeError(f"{total} points exceed the engine capacity {self.cap}")
        # Frames that already live in pinned host memory (what a pinning DataLoader hands over) go to the device slot
        # directly, one cudaMemcpyAsync per frame: the extra host->host copy into the engine's own pinned slot cost
        # more than the H2D itself once 8 ranks shared one socket's memory bandwidth (0.89 e2e scaling at N=8, r1)
        direct = all(f.is_pinned() and f.is_contiguous() for f in frames)
        n, ho[0] = 0, 0
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._ev_used[slot])    # the previous batch in this slot has been consumed
            for b, f in enumerate(frames):
                k = f.shape[0]
                if direct:
                    dp[n:n + k].copy_(f, non_blocking=True)
                else:
                    hp[n:n + k].copy_(f)
                n += k
                ho[b + 1] = n
            if not direct:
                dp[:n].copy_(hp[:n], non_blocking=True)
            do.copy_(ho, non_blocking=True)
            self._ev_up[slot].record(self._copy_stream)
        self._n[slot] = n
        return n * dp.shape[1] * 4 + do.numel() * 4

    def _launch_slot(self, slot):
        """queue the batch staged in input slot `slot` on the engine stream: refresh the graph's input buffers from the
        slot, replay, read the result back into the slot's pinned buffers; returns the bytes read back"""
        dp, do = self._p_dev[slot]
        rd, rc = self._p_res[slot]
        n = self._n[slot]
        with torch.cuda.stream(self.stream), torch.no_grad():
            self.stream.wait_event(self._ev_up[slot])
            self.points[:n].copy_(dp[:n], non_blocking=True)
            self.offsets.copy_(do, non_blocking=True)
            self._ev_used[slot].record(self.stream)
            if self.graph is not None:
                self.graph.replay()
            else:
                self.det_out, self.keep_count, self.plan = self._forward()
            rd.copy_(self.det_out, non_blocking=True)
            rc.copy_(self.keep_count, non_blocking=True)
            self._ev_done[slot].record(self.stream)
        return rd.numel() * 4 + rc.numel() * 4

    def run_pipelined(self, batches, consume=None):
        """Runs a sequence of batches (each a list of B frames) with the next batch's host packing and H2D
        overlapped with the current batch's graph replay.  `consume(i, detections)` is called per batch (default:
        collect and return).  Returns (results, h2d_bytes_per_batch, d2h_bytes_per_batch)."""
        if self.det_out is None and len(batches) > 0:
            self.upload(self.stage_host(batches[0]))    # the warm-up passes see a real batch
            self.prepare()
        elif self.det_out is None:
            self.prepare()
        if not hasattr(self, "_p_host"):
            self._init_pipeline()
        out, h2d, d2h = [], 0, 0
        n_b = len(batches)
        if n_b == 0:
            return out, 0, 0
        def launch(i):
            return self._launch_slot(i & 1)

        h2d = self._prefetch(0, batches[0])
        d2h = launch(0)
        for i in range(n_b):
            slot = i & 1
            if i + 1 < n_b:
                # batch i is running: pack + upload batch i+1 and queue it behind, so the device never idles while
                # the host assembles batch i's detections
                h2d = self._prefetch(slot ^ 1, batches[i + 1])
                launch(i + 1)
            self._ev_done[slot].synchronize()
            self.h_det, self.h_cnt = self._p_res[slot]
            dets = self.assemble_host()
            if consume is not None:
                consume(i, dets)
            else:
                out.append(dets)
        return out, h2d, d2h

    def assemble_host(self, metadata=None):
        """detections from the pinned read-back, det3d's structure (center_head.py:332-350,405-409)"""
        head = self.model.bbox_head
        plan = self.plan
        B, S, post_cap = plan["B"], plan["S"], plan["post_cap"]
        det = self.h_det.view(B, S, post_cap, 11)
        cnt = self.h_cnt.view(B, S)
        cls_off, flag = [], 0
        for k in head.num_classes:
            cls_off.append(flag)
            flag += k
        out = []
        for b in range(B):
            boxes, scores, labels = [], [], []
            for s, seg in enumerate(plan["segs"]):
                k = int(cnt[b, s])
                d = det[b, s, :k]
                boxes.append(d[:, :9] if head.box_n_dim == 9 else torch.cat([d[:, :6], d[:, 8:9]], 1))
                scores.append(d[:, 9])
                labels.append(d[:, 10].to(torch.int64) + cls_off[seg["task"]])
            out.append({"box3d_lidar": torch.cat(boxes), "scores": torch.cat(scores),
                        "label_preds": torch.cat(labels), "metadata": metadata[b] if metadata else None})
        return out


class StreamingEngine:
    """Several batches in flight on one GPU: `in_flight` lanes, each an InferenceEngine with its own stream, CUDA graph
    and buffers, take the batches round-robin.  A batch-1 step ends in kernels that occupy a handful of SMs (top-k
    selection, NMS mask and sweep: ~90 us on <= 40 CTAs) and starts with small reader / rulebook launches; with
    another frame's convs running beside them those phases cost nothing (measured on one B200, nuScenes batch 1:
    1113 -> 1236 -> 1300 frames/s with 1 / 2 / 3 frames in flight).  Per-frame results are identical to the single
    lane's (tests/test_gpu_model.py); per-frame latency grows to about `in_flight` steps."""

    def __init__(self, model, n_frames, points_cap, in_flight=3, point_dim=5, device=None):
        assert in_flight >= 1
        self.lanes = [InferenceEngine(model, n_frames, points_cap, point_dim, device) for _ in range(in_flight)]
        self.model, self.B, self.dev = self.lanes[0].model, n_frames, self.lanes[0].dev
        self._fork = torch.cuda.Event(enable_timing=True)
        self._join = torch.cuda.Event(enable_timing=True)

    def prepare(self, frames, warmup=2):
        """warm-up and graph capture of every lane on a real batch"""
        for lane in self.lanes:
            if lane.det_out is None:
                lane.upload(lane.stage_host(frames))
                lane.prepare(warmup)
        return self

    @property
    def plan(self):
        return self.lanes[0].plan

    def launch_resident(self, i, points, offsets):
        """step i from device-resident inputs on lane i % in_flight (no host synchronisation)"""
        lane = self.lanes[i % len(self.lanes)]
        with torch.cuda.stream(lane.stream):
            lane.points[:points.shape[0]].copy_(points, non_blocking=True)
            lane.offsets.copy_(offsets, non_blocking=True)
        lane.launch()
        return lane

    def fork(self, stream=None):
        """start of a timed region: every lane waits for this point of `stream` (default: the current stream)"""
        stream = stream or torch.cuda.current_stream(self.dev)
        self._fork.record(stream)
        for lane in self.lanes:
            lane.stream.wait_event(self._fork)
        return self._fork

    def join(self, stream=None):
        """end of a timed region: `stream` waits for every lane; returns (fork event, join event) for elapsed_time"""
        stream = stream or torch.cuda.current_stream(self.dev)
        for lane in self.lanes:
            ev = torch.cuda.Event()
            ev.record(lane.stream)
            stream.wait_event(ev)
        self._join.record(stream)
        return self._fork, self._join

    def synchronize(self):
        for lane in self.lanes:
            lane.stream.synchronize()

    def run(self, batches, consume=None):
        """Throughput API over host batches (each a list of B frames, ideally in pinned memory): batch i goes to lane
        i % in_flight; per lane the next batch's packing + H2D overlap the current replay (two input / result slots
        per lane), so up to 2 * in_flight batches are queued.  `consume(i, detections)` is called in batch order
        (default: collect).  Returns (results, h2d_bytes_per_batch, d2h_bytes_per_batch)."""
        n_b, L = len(batches), len(self.lanes)
        if n_b == 0:
            return [], 0, 0
        self.prepare(batches[0])
        for lane in self.lanes:
            if not hasattr(lane, "_p_host"):
                lane._init_pipeline()
        depth = 2 * L
        h2d = d2h = 0

        def issue(i):
            lane, slot = self.lanes[i % L], (i // L) & 1
            a = lane._prefetch(slot, batches[i])
            return a, lane._launch_slot(slot)

        for i in range(min(depth, n_b)):
            h2d, d2h = issue(i)
        out = []
        for i in range(n_b):
            lane, slot = self.lanes[i % L], (i // L) & 1
            lane._ev_done[slot].synchronize()
            lane.h_det, lane.h_cnt = lane._p_res[slot]
            dets = lane.assemble_host()
            if consume is not None:
                consume(i, dets)
            else:
                out.append(dets)
            if i + depth < n_b:
                h2d, d2h = issue(i + depth)
        return out, h2d, d2h


def calibrate_heatmap_bias(model, frames, target_cells=1500, calibrate_rot=True):
    """Random-init heads give a degenerate candidate count (hm bias -2.19, center_head.py:19,38): shift each
    task's heat-map bias so that about `target_cells` cells per frame pass the score threshold, which is
    what a trained model feeds the NMS stage."""
    import math
    dev = next(model.parameters()).device
    thr = float(model.test_cfg["score_threshold"])
    logit_thr = math.log(thr / (1 - thr))
    counts = np.cumsum([0] + [len(f) for f in frames]).astype(np.int32)
    pts = torch.from_numpy(np.concatenate(frames)).to(dev)
    off = torch.from_numpy(counts).to(dev)
    with torch.no_grad():
        bev, _ = model.extract_feat(dict(points_batched=(pts, off)))
        preds = model.bbox_head(bev)
        for t, p in enumerate(preds):
            hm = p["hm"].float().amax(1).reshape(len(frames), -1)
            k = max(1, min(hm.shape[1] - 1, target_cells))
            kth = torch.topk(hm, k, dim=1).values[:, -1].mean().item()
            fc = model.bbox_head.task_heads[t].hm
            fc[-1].bias.add_(logit_thr - kth)   # in-place on the Parameter: bumps ._version -> lowering refresh
            # A trained rot head outputs (sin, cos) of unit norm; a random-init one outputs two numbers near zero, and
            # atan2 of those turns a 1e-2 perturbation into an arbitrary heading (center_head.py:270-271).  Bias the
            # cos channel so the decoded boxes are conditioned like a trained model's.
            rot = getattr(model.bbox_head.task_heads[t], "rot", None)
            if rot is not None and calibrate_rot:
                rot[-1].bias.copy_(torch.tensor([0.0, 1.0], device=rot[-1].bias.device))
    torch.cuda.synchronize()


def pin_process_to_gpu_cores(local_rank, world_local):
    """Binds this process to host cores next to its GPU: the NVML CPU-affinity set of device `local_rank`, split
    evenly among the local ranks that share that set (one process per GPU; with 8 ranks on a two-socket host every
    rank otherwise floats over all cores and packs its frames through the remote socket's memory).  Returns the list
    of cores, or None when NVML / sched_setaffinity is unavailable (then nothing is changed)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        n_cpu = os.cpu_count() or 1
        words = (n_cpu + 63) // 64

        def cores_of(idx):
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            return tuple(c for c in range(n_cpu) if (mask[c // 64] >> (c % 64)) & 1)

        allowed = set(os.sched_getaffinity(0))
        mine = cores_of(local_rank)
        sharers = [r for r in range(world_local) if cores_of(r) == mine]
        mine = [c for c in mine if c in allowed]
        if not mine:
            return None
        per = max(1, len(mine) // max(1, len(sharers)))
        k = sharers.index(local_rank)
        sel = mine[k * per:(k + 1) * per] or mine
        os.sched_setaffinity(0, sel)
        return list(sel)
    except Exception:
        return None
