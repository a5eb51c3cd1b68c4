"""The tracker's per-frame step, device-resident and in the reference's order (tools/tracking/demo.py:35-42 process_frame):

    flow       = flow_est(flownet, prev_im, cur_im)                       lib/tracking/net_utils.py:73-92
    prop_boxes = box_propagation(prev_keypoints, flow)                    lib/tracking/flow_utils.py:7-35
    prop_dets  = [prop_boxes | prev_dets[:, 4]]                           demo.py:38
    cur_dets   = detect(...): NMS over [detector boxes | prop_dets]       net_utils.py:25-33 (cpu_nms rule: IoU >= thresh)
    cur_kps    = pose_est(pose_net, cur_im, cur_dets[:, :4])              net_utils.py:36-71 (crop -> PoseResNet -> final_preds)

for `streams` independent videos at once (one frame pair each per step) with up to `people` person slots per frame.  The
pose crops DEPEND on the flow (they are cut at the propagated, NMS-filtered boxes), so inside one step the chain is serial;
what overlaps is the flow network of step t+1 with the box/pose tail of step t -- flow(t+1) needs only frames, never the
key-points -- through two lanes of input/flow buffers and two CUDA graphs per lane:

    copy stream : frames(t+1) host -> device (uint8 BGR, 6 bytes per pixel pair; detector boxes)
    flow stream : GF[lane]  = FlowNet2* forward straight from the uint8 frames (ft_pack_pair_u8)
    tail stream : GT[lane]  = ft_box_propagation -> ft_track_merge -> ft_nms -> ft_track_select -> ft_crop_affine (from the
                              uint8 current frame) -> PoseResNet -> ft_decode_keypoints; results -> pinned host

The recurrent state (person boxes + scores, their count, the key-points) stays on the GPU between steps; every step also
returns it to the host.  Detection (Faster R-CNN) is out of scope: detector boxes are an optional input of each step
(`max_det` per frame), exactly where process_frame merges them.

Multi-GPU: videos are independent; each rank owns `shard_pairs(num_streams, rank, world)` of them, runs this
pipeline on its shard, and `gather(lane)` all-gathers the step's packed per-person records once per step (one NCCL
all-gather on its own stream behind the tail).
"""
import torch

from .. import _lib, engine
from . import flow_utils, net_utils


class _Lane:
    pass


class TrackingPipeline:
    def __init__(self, pose_net, flow_net, streams, people, frame_hw=(384, 512), crop_hw=(256, 192), device="cuda", max_det=0, nms_thresh=0.3,
                 use_cuda_graph=True, lanes=2):
        self.pose_net, self.flow_net = pose_net, flow_net
        self.F, self.P, self.D, self.n = streams, people, max_det, streams * people
        self.N = self.D + self.P                                            # NMS candidates per frame
        self.device = dev = torch.device(device)
        self.frame_hw, self.crop_hw, self.nms_thresh = tuple(frame_hw), tuple(crop_hw), float(nms_thresh)
        self.K = pose_net.num_classes
        self.use_cuda_graph = use_cuda_graph
        F, P, N, K, n = self.F, self.P, self.N, self.K, self.n
        H, W = self.frame_hw
        with torch.cuda.device(dev):
            f32, i32 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int32, device=dev)
            # recurrent state
            self.dets, self.count = torch.zeros(F, P, 5, **f32), torch.zeros(F, **i32)
            self.keypoints = torch.zeros(n, K, 3, **f32)
            # scratch of the tail
            self.prop_boxes, self.cand, self.cand_count = torch.zeros(n, 4, **f32), torch.zeros(F, N, 5, **f32), torch.zeros(F, **i32)
            self.keep, self.num_keep = torch.zeros(F, N, **i32), torch.zeros(F, **i32)
            self.centers, self.scales, self.valid = torch.zeros(n, 2, **f32), torch.ones(n, **f32), torch.zeros(n, **i32)
            self.frame_of = torch.arange(n, **i32) // P
            self.pose_plan, self.crops, self.heatmaps = pose_net.plan_for(n, crop_hw[0], crop_hw[1], dev)
            self.n_out = F * P * 5 + F + n * K * 3                          # dets | count (as float) | key-points
            self.lanes = []
            for l in range(lanes):
                ln = _Lane()
                ln.flow_plan, ln.frames, ln.flow = flow_net.plan_for(F, H, W, dev, u8=True, tag=("track", l))
                ln.det, ln.det_count = torch.zeros(F, max(self.D, 1), 5, **f32), torch.zeros(F, **i32)
                ln.h_frames = torch.empty(ln.frames.shape, dtype=torch.uint8, pin_memory=True)
                ln.h_det = torch.zeros(ln.det.shape, dtype=torch.float32, pin_memory=True)
                ln.h_det_count = torch.zeros(F, dtype=torch.int32, pin_memory=True)
                ln.d_out, ln.h_out = torch.zeros(self.n_out, **f32), torch.zeros(self.n_out, dtype=torch.float32, pin_memory=True)
                ln.copied, ln.flow_done, ln.done, ln.gathered = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
                ln.g_out = None
                ln.gf = ln.gt = None
                ln.busy = False
                self.lanes.append(ln)
            self.s_copy, self.s_flow, self.s_tail = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            self.s_comm = torch.cuda.Stream(device=dev)
        self._next = 0
        self.launches_per_step = 0

    # ---- state -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def seed(self, dets, count, keypoints):
        """Set the recurrent state: dets [F,P,5] (x1,y1,x2,y2,score), count [F] people per frame, keypoints [F*P,K,3]."""
        with torch.cuda.device(self.device), torch.cuda.stream(self.s_tail):
            self.dets.copy_(torch.as_tensor(dets, dtype=torch.float32).reshape(self.F, self.P, 5), non_blocking=False)
            self.count.copy_(torch.as_tensor(count, dtype=torch.int32).reshape(self.F))
            self.keypoints.copy_(torch.as_tensor(keypoints, dtype=torch.float32).reshape(self.n, self.K, 3))
        self.s_tail.synchronize()

    # ---- the two halves of a step ------------------------------------------------------------------------
    def _enqueue_flow(self, ln):
        ln.flow_plan.enqueue()

    def _enqueue_tail(self, ln):
        L, s = _lib.lib(), engine.stream_ptr
        F, P, D, N, n = self.F, self.P, self.D, self.N, self.n
        H, W = self.frame_hw
        flow_utils.box_propagation_cuda(self.keypoints, ln.flow, self.frame_of, out=self.prop_boxes)
        _lib.check(L.ft_track_merge(self.prop_boxes.data_ptr(), self.dets.data_ptr(), self.count.data_ptr(), ln.det.data_ptr() if D else None,
                                    ln.det_count.data_ptr() if D else None, F, P, D, self.cand.data_ptr(), self.cand_count.data_ptr(), s()), "track_merge")
        _lib.check(L.ft_nms(self.cand.data_ptr(), self.cand_count.data_ptr(), F, N, self.nms_thresh, 1, self.keep.data_ptr(), self.num_keep.data_ptr(), s()), "nms")
        _lib.check(L.ft_track_select(self.cand.data_ptr(), self.keep.data_ptr(), self.num_keep.data_ptr(), F, N, P, self.crop_hw[0], self.crop_hw[1], H, W,
                                     self.dets.data_ptr(), self.count.data_ptr(), self.centers.data_ptr(), self.scales.data_ptr(), self.valid.data_ptr(), s()),
                   "track_select")
        # person crops out of the uint8 BGR current frame of every pair (frames[:, 1]), as pose_est cuts them (net_utils.py:52-57)
        xs = self.pose_plan.taps.get("input_s2d")
        if xs is not None and self.pose_plan.calls[0][0] == "pack_image_s2d":
            # ... written straight into the stem's space-to-depth split-fp16 input (ft_crop_affine_s2d): the fp32 crops never exist
            fr = ln.frames[:, 1]
            sn, sh, sw, sc = fr.stride()
            _lib.check(L.ft_crop_affine_s2d(fr.data_ptr(), 1, sn, sc, sh, sw, H, W, 3, self.centers.data_ptr(), self.scales.data_ptr(), self.frame_of.data_ptr(),
                                            n, self.crop_hw[0], self.crop_hw[1], 0, xs.struct(), xs.xl, xs.xr, s()), "crop_affine_s2d")
            self.pose_plan.enqueue(start=1)
        else:
            net_utils.crop_boxes_cuda(ln.frames[:, 1], self.centers, self.scales, self.frame_of, self.crop_hw, reverse_channels=False, out=self.crops)
            self.pose_plan.enqueue()
        net_utils.final_preds_cuda(self.heatmaps, self.centers, self.scales, True, out=self.keypoints, valid=self.valid)
        o, a, b = ln.d_out, F * P * 5, F * P * 5 + F
        o[:a].copy_(self.dets.view(-1))
        o[a:b].copy_(self.count)
        o[b:].copy_(self.keypoints.view(-1))

    def person_crops(self, lane):
        """The fp32 crops [F*P,3,h,w] of the lane's last step, recomputed from the state it left (centres / scales of the selected
        boxes, the lane's current frames) -- the fused path never materialises them; tests and the bench's parity leg use this."""
        ln = self.lanes[lane]
        return net_utils.crop_boxes_cuda(ln.frames[:, 1], self.centers, self.scales, self.frame_of, self.crop_hw, reverse_channels=False)

    def capture(self):
        with torch.cuda.device(self.device):
            saved = (self.dets.clone(), self.count.clone(), self.keypoints.clone())
            for ln in self.lanes:                                            # warm-up: function attributes, code upload
                self._enqueue_flow(ln)
                self._enqueue_tail(ln)
            torch.cuda.synchronize()
            _lib.reset_launch_count()
            for i, ln in enumerate(self.lanes):
                ln.gf, ln.gt = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(ln.gf):
                    self._enqueue_flow(ln)
                with torch.cuda.graph(ln.gt):
                    self._enqueue_tail(ln)
                if i == 0:
                    self.launches_per_step = _lib.launch_count()
            torch.cuda.synchronize()
            for dst, src in zip((self.dets, self.count, self.keypoints), saved):   # the warm-up steps advanced the state: put it back
                dst.copy_(src)
            torch.cuda.synchronize()
        return self

    def _run_flow(self, ln):
        if self.use_cuda_graph:
            ln.gf.replay()
        else:
            self._enqueue_flow(ln)

    def _run_tail(self, ln):
        if self.use_cuda_graph:
            ln.gt.replay()
        else:
            self._enqueue_tail(ln)

    def _ready(self):
        if self.use_cuda_graph and self.lanes[0].gf is None:
            self.capture()
        elif not self.use_cuda_graph and not self.launches_per_step:
            _lib.reset_launch_count()
            with torch.cuda.device(self.device):
                saved = (self.dets.clone(), self.count.clone(), self.keypoints.clone())
                self._enqueue_flow(self.lanes[0])
                self._enqueue_tail(self.lanes[0])
                torch.cuda.synchronize()
                for dst, src in zip((self.dets, self.count, self.keypoints), saved):
                    dst.copy_(src)
                torch.cuda.synchronize()
            self.launches_per_step = _lib.launch_count()

    # ---- device-resident step (inputs already in the lane's device buffers) ------------------------------
    @torch.no_grad()
    def step_device(self, lane=None):
        """One step on the frames / detector boxes already resident in the lane's device buffers: flow on the flow stream,
        tail on the tail stream (ordered after this lane's flow and after the previous step's tail).  No host sync."""
        self._ready()
        l = self._next if lane is None else lane
        ln = self.lanes[l]
        with torch.cuda.device(self.device):
            with torch.cuda.stream(self.s_flow):
                self.s_flow.wait_event(ln.done)                          # the lane's previous tail has read its flow / frames
                self._run_flow(ln)
                ln.flow_done.record(self.s_flow)
            with torch.cuda.stream(self.s_tail):
                self.s_tail.wait_event(ln.flow_done)
                self.s_tail.wait_event(ln.gathered)                      # the lane's previous records have left for the other ranks
                self._run_tail(ln)
                ln.done.record(self.s_tail)
        self._next = (l + 1) % len(self.lanes)
        return l

    @torch.no_grad()
    def gather(self, lane, group=None):
        """The single collective of the path: all-gather of the step's packed records (dets | count | key-points of
        every person slot, `n_out` floats per rank) over NCCL, on its own stream behind the lane's tail -- it neither delays the
        next step's kernels nor waits for them.  Returns the [world, n_out] device tensor (valid until the lane's next step), or
        the local records as [1, n_out] when no process group is initialised."""
        import torch.distributed as dist
        ln = self.lanes[lane]
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return ln.d_out.view(1, -1)
        world = dist.get_world_size(group)
        with torch.cuda.device(self.device), torch.cuda.stream(self.s_comm):
            if ln.g_out is None:
                ln.g_out = torch.empty(world, self.n_out, dtype=torch.float32, device=self.device)
            self.s_comm.wait_event(ln.done)
            dist.all_gather_into_tensor(ln.g_out.view(-1), ln.d_out, group=group)
            ln.gathered.record(self.s_comm)
        return ln.g_out

    def fence_in(self):
        """The pipeline's streams wait for everything enqueued so far on the caller's current stream."""
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_copy, self.s_flow, self.s_tail, self.s_comm):
            s.wait_stream(cur)

    def fence_out(self):
        """The caller's current stream waits for everything the pipeline has enqueued."""
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_copy, self.s_flow, self.s_tail, self.s_comm):
            cur.wait_stream(s)

    # ---- host-facing step --------------------------------------------------------------------------------
    @property
    def h2d_bytes(self):
        ln = self.lanes[0]
        return ln.h_frames.numel() + (4 * ln.h_det.numel() + 4 * ln.h_det_count.numel() if self.D else 0)

    @property
    def d2h_bytes(self):
        return 4 * self.n_out

    def stage(self, frames, det=None, det_count=None, lane=None):
        """Fill a lane's pinned host buffers: frames uint8 [F,2,H,W,3] (prev, cur; cv2 BGR), optional detector boxes
        det [F,D,5] + det_count [F].  Callers may also write into `lanes[l].h_frames` directly (it is the pinned buffer the
        copy engine reads) and skip this memcpy.  Returns the lane index."""
        l = self._next if lane is None else lane
        ln = self.lanes[l]
        if ln.busy:
            raise RuntimeError("lane %d still holds an uncollected step: collect() it first" % l)
        ln.h_frames.copy_(torch.as_tensor(frames).reshape(ln.h_frames.shape))
        if self.D:
            if det is None:
                ln.h_det_count.zero_()
            else:
                ln.h_det.copy_(torch.as_tensor(det, dtype=torch.float32).reshape(ln.h_det.shape))
                ln.h_det_count.copy_(torch.as_tensor(det_count, dtype=torch.int32).reshape(self.F))
        return l

    @torch.no_grad()
    def submit(self, lane=None):
        """Enqueue one step from a staged lane: pinned host -> the lane's device buffers (copy stream), flow (flow stream), tail
        (tail stream), results -> the lane's pinned output.  Nothing blocks the host."""
        self._ready()
        l = self._next if lane is None else lane
        ln = self.lanes[l]
        if ln.busy:
            raise RuntimeError("lane %d still holds an uncollected step: collect() it first" % l)
        with torch.cuda.device(self.device):
            with torch.cuda.stream(self.s_copy):
                self.s_copy.wait_event(ln.done)                          # the lane's previous step no longer reads its frames
                ln.frames.copy_(ln.h_frames, non_blocking=True)
                if self.D:
                    ln.det.copy_(ln.h_det, non_blocking=True)
                    ln.det_count.copy_(ln.h_det_count, non_blocking=True)
                ln.copied.record(self.s_copy)
            with torch.cuda.stream(self.s_flow):
                self.s_flow.wait_event(ln.copied)
                self._run_flow(ln)
                ln.flow_done.record(self.s_flow)
            with torch.cuda.stream(self.s_tail):
                self.s_tail.wait_event(ln.flow_done)
                self.s_tail.wait_event(ln.gathered)
                self._run_tail(ln)
                ln.h_out.copy_(ln.d_out, non_blocking=True)
                ln.done.record(self.s_tail)
        ln.busy = True
        self._next = (l + 1) % len(self.lanes)
        return l

    def collect(self, lane):
        """Wait for a submitted step; returns (dets [F,P,5], count [F] int, keypoints [F*P,K,3]) as views / copies of the lane's
        pinned output (valid until the lane is submitted again)."""
        ln = self.lanes[lane]
        if not ln.busy:
            raise RuntimeError("lane %d has no step in flight" % lane)
        ln.done.synchronize()
        ln.busy = False
        F, P, K, n = self.F, self.P, self.K, self.n
        a, b = F * P * 5, F * P * 5 + F
        return ln.h_out[:a].view(F, P, 5), ln.h_out[a:b].round().int(), ln.h_out[b:].view(n, K, 3)

    @torch.no_grad()
    def process(self, frames, det=None, det_count=None):
        """Host frames in, host results out (one synchronous step): stage -> submit -> collect."""
        return self.collect(self.submit(self.stage(frames, det, det_count)))
