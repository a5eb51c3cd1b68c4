"""The joint per-frame-pair step (the intended `process_frame` of tools/tracking/demo.py:35-42, minus detection):

    flow      = flow_net(pair)                      FlowNet2-family forward          (net_utils.flow_est, :73-92)
    boxes     = box_propagation(prev_kps, flow)     flow look-up at the key-points   (flow_utils.py:7-35)
    heatmaps  = pose_net(crops)                     PoseResNet forward               (net_utils.pose_est, :36-71)
    keypoints = final_preds(heatmaps, centers, scales, adjust_coords=True)           (evaluation.py:22-35)

for a batch of `pairs` frame pairs with `people` person crops each, as ONE CUDA graph: the flow net and the pose
net are captured on two forked streams (they are independent until the box propagation), the decode and the
propagation kernels join them.  All buffers are static, so a step is: copy inputs in, replay, copy results out.

Multi-GPU: frame pairs are independent, so each rank owns a shard of pairs and runs this
pipeline on it; `gather_keypoints` is the single collective of the path -- one NCCL all-gather of the per-person
key-point records (and boxes) at the end of a batch.  `shard_pairs` is the partition rule (also used by the gloo
CPU tests)."""
import torch

from .. import _lib, engine
from . import flow_utils, net_utils


def shard_pairs(num_pairs, rank, world):
    """Contiguous block partition of pair indices: rank r gets [lo, hi).  Sizes differ by at most one."""
    base, rem = divmod(num_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_keypoints(local, group=None):
    """All-gather equally sized per-rank record tensors [n_local, ...] -> [world*n_local, ...] (rank-major).
    Works on CUDA tensors over NCCL and on CPU tensors over gloo."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    else:
        dist.all_gather(list(out.chunk(world)), local.contiguous(), group=group)
    return out


class FramePairPipeline:
    def __init__(self, pose_net, flow_net, pairs, people, frame_hw=(384, 512), crop_hw=(256, 192), device="cuda", use_cuda_graph=True,
                 concurrent=True, crops_from_frames=False):
        """crops_from_frames: cut the person crops on the GPU (ft_crop_affine, the cv2.warpAffine rule of pose_est,
        net_utils.py:48-57) out of the current frame of every pair -- the frame is already on the device for the flow net --
        instead of taking ready-made crops from the host.  The crops are the BGR view of the RGB flow input, raw 0..255
        values, as the reference feeds them."""
        self.pose_net, self.flow_net = pose_net, flow_net
        self.crops_from_frames = crops_from_frames
        self.pairs, self.people, self.n = pairs, people, pairs * people
        self.device = torch.device(device)
        self.frame_hw, self.crop_hw = frame_hw, crop_hw
        self.num_joints = pose_net.num_classes
        dev = self.device
        self.pose_plan, self.crops, self.heatmaps = pose_net.plan_for(self.n, crop_hw[0], crop_hw[1], dev)
        self.flow_plan, self.pair, self.flow = flow_net.plan_for(pairs, frame_hw[0], frame_hw[1], dev)
        f32 = dict(dtype=torch.float32, device=dev)
        self.centers, self.scales = torch.zeros(self.n, 2, **f32), torch.ones(self.n, **f32)
        self.prev_kps = torch.zeros(self.n, self.num_joints, 3, **f32)
        self.frame_of = torch.arange(self.n, dtype=torch.int32, device=dev) // people      # person i belongs to pair i // people
        self.keypoints = torch.zeros(self.n, self.num_joints, 3, **f32)
        self.boxes = torch.zeros(self.n, 4, **f32)
        self.use_cuda_graph, self.concurrent = use_cuda_graph, concurrent
        self.graph, self.launches_per_step = None, 0
        self._side = torch.cuda.Stream(device=dev)
        # host-facing call: two slots of pinned host staging + device staging, so the host->device copy of batch i+1
        # (copy stream) overlaps the graph replay of batch i (main stream); see submit() / collect()
        pin = dict(dtype=torch.float32, pin_memory=True)
        n_meta, n_out = self.n * (2 + 1 + self.num_joints * 3), self.n * (self.num_joints * 3 + 4)
        self._slots = []
        crop_shape = (0,) if crops_from_frames else tuple(self.crops.shape)       # no crop staging when they are cut on the device
        for _ in range(2):
            self._slots.append(dict(
                h_pair=torch.empty(self.pair.shape, **pin), h_crops=torch.empty(crop_shape, **pin), h_meta=torch.empty(n_meta, **pin),
                h_out=torch.empty(n_out, **pin),
                d_pair=torch.empty(self.pair.shape, **f32), d_crops=torch.empty(crop_shape, **f32), d_meta=torch.empty(n_meta, **f32),
                d_out=torch.empty(n_out, **f32),
                copied=torch.cuda.Event(), done=torch.cuda.Event(), consumed=torch.cuda.Event(), busy=False))
        self._copy = torch.cuda.Stream(device=dev)
        self._next = 0

    # ---- device-side step ------------------------------------------------------------------------------
    def _enqueue(self):
        main = torch.cuda.current_stream()
        if self.concurrent:
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                self.flow_plan.enqueue()
        else:
            self.flow_plan.enqueue()
        if self.crops_from_frames:
            net_utils.crop_boxes_cuda(self.pair[:, :, 1], self.centers, self.scales, self.frame_of, self.crop_hw, reverse_channels=True, out=self.crops)
        self.pose_plan.enqueue()
        net_utils.final_preds_cuda(self.heatmaps, self.centers, self.scales, True, out=self.keypoints)
        if self.concurrent:
            main.wait_stream(self._side)
        flow_utils.box_propagation_cuda(self.prev_kps, self.flow, self.frame_of, out=self.boxes)

    def capture(self):
        with torch.cuda.device(self.device):
            self._enqueue()                      # warm-up: function attributes, code upload
            torch.cuda.synchronize()
            _lib.reset_launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()
            self.launches_per_step = _lib.launch_count()
            self.graph = g
        return self

    @torch.no_grad()
    def step_device(self):
        """One batch with inputs already resident in the static device buffers (self.pair, self.crops, self.centers,
        self.scales, self.prev_kps).  Results land in self.keypoints / self.boxes / self.flow.  No host sync."""
        if self.use_cuda_graph:
            if self.graph is None:
                self.capture()
            self.graph.replay()
        else:
            if not self.launches_per_step:
                _lib.reset_launch_count()
                self._enqueue()
                self.launches_per_step = _lib.launch_count()
            else:
                self._enqueue()

    # ---- host-facing step --------------------------------------------------------------------------------
    @property
    def h2d_bytes(self):
        sl = self._slots[0]
        return 4 * (sl["h_pair"].numel() + sl["h_crops"].numel() + sl["h_meta"].numel())

    @property
    def d2h_bytes(self):
        return 4 * self._slots[0]["h_out"].numel()

    def stage(self, pairs, crops, centers, scales, prev_keypoints, slot=None):
        """Fill the pinned host staging buffers of a slot (the next one by default) with one batch; returns the slot index.
        pairs [pairs,3,2,H,W] (RGB 0..255), crops [n,3,h,w] (ignored / None with crops_from_frames), centers [n,2],
        scales [n], prev_keypoints [n,K,3] (fp32)."""
        k = self._next if slot is None else slot
        sl, n = self._slots[k], self.n
        if sl["busy"]:
            raise RuntimeError("slot %d still holds an uncollected batch" % k)
        sl["h_pair"].copy_(pairs)
        if not self.crops_from_frames:
            sl["h_crops"].copy_(crops)
        m = sl["h_meta"]
        m[:2 * n].copy_(torch.as_tensor(centers, dtype=torch.float32).reshape(-1))
        m[2 * n:3 * n].copy_(torch.as_tensor(scales, dtype=torch.float32).reshape(-1))
        m[3 * n:].copy_(torch.as_tensor(prev_keypoints, dtype=torch.float32).reshape(-1))
        return k

    @torch.no_grad()
    def submit(self, slot=None):
        """Enqueue one batch from a staged slot: pinned host -> device on the copy stream, then (main stream) move it into
        the graph's static inputs, replay, and send the results back to the slot's pinned output.  Returns the slot index;
        nothing blocks the host, so the copy-in of the next submit() overlaps this batch's kernels."""
        k = self._next if slot is None else slot
        sl, n, K = self._slots[k], self.n, self.num_joints
        if sl["busy"]:
            raise RuntimeError("slot %d still holds an uncollected batch: collect() it first" % k)
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self._copy):
            self._copy.wait_event(sl["consumed"])            # the previous batch of this slot has left the device staging
            sl["d_pair"].copy_(sl["h_pair"], non_blocking=True)
            if not self.crops_from_frames:
                sl["d_crops"].copy_(sl["h_crops"], non_blocking=True)
            sl["d_meta"].copy_(sl["h_meta"], non_blocking=True)
            sl["copied"].record(self._copy)
        main.wait_event(sl["copied"])
        self.pair.copy_(sl["d_pair"])
        if not self.crops_from_frames:
            self.crops.copy_(sl["d_crops"])
        self.centers.copy_(sl["d_meta"][:2 * n].view(n, 2))
        self.scales.copy_(sl["d_meta"][2 * n:3 * n])
        self.prev_kps.copy_(sl["d_meta"][3 * n:].view(n, K, 3))
        sl["consumed"].record(main)
        self.step_device()
        sl["d_out"][:n * K * 3].copy_(self.keypoints.view(-1))
        sl["d_out"][n * K * 3:].copy_(self.boxes.view(-1))
        sl["h_out"].copy_(sl["d_out"], non_blocking=True)
        sl["done"].record(main)
        sl["busy"] = True
        self._next = k ^ 1
        return k

    def collect(self, slot):
        """Wait for a submitted batch; returns (keypoints [n,K,3], boxes [n,4]) as views of the slot's pinned host buffer
        (valid until the slot is staged again)."""
        sl, n, K = self._slots[slot], self.n, self.num_joints
        if not sl["busy"]:
            raise RuntimeError("slot %d has no batch in flight" % slot)
        sl["done"].synchronize()
        sl["busy"] = False
        return sl["h_out"][:n * K * 3].view(n, K, 3), sl["h_out"][n * K * 3:].view(n, 4)

    @torch.no_grad()
    def process(self, pairs, crops, centers, scales, prev_keypoints):
        """Host tensors in, host tensors out (one synchronous batch): stage -> submit -> collect."""
        return self.collect(self.submit(self.stage(pairs, crops, centers, scales, prev_keypoints)))

    @torch.no_grad()
    def process_staged(self, slot=None):
        """The synchronous timed part of process(): pinned host staging buffers -> device, one step, results -> pinned host."""
        return self.collect(self.submit(slot))
