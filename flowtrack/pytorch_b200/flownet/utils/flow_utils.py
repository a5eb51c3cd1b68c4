"""Middlebury .flo files with the reference names (lib/flownet/utils/flow_utils.py:5-55): the on-disk
format of the step after the flow forward (tools/flownet/main.py --save_flow, demo.py).

Layout (little endian): float32 magic 202021.25 | int32 width | int32 height | height*width interleaved (u, v) float32.
Host I/O only; `writeFlow` also accepts the network's CUDA output ([2,H,W] or [H,W,2]) and copies it once."""
import numpy as np

TAG_CHAR = np.array([202021.25], np.float32)


def readFlow(fn):
    """-> ndarray [H, W, 2] float32, or None when the magic number is wrong (the reference prints and returns None)."""
    with open(fn, "rb") as f:
        magic = np.fromfile(f, np.float32, count=1)
        if magic.size != 1 or magic[0] != TAG_CHAR[0]:
            print("Magic number incorrect. Invalid .flo file")
            return None
        w = int(np.fromfile(f, np.int32, count=1)[0])
        h = int(np.fromfile(f, np.int32, count=1)[0])
        data = np.fromfile(f, np.float32, count=2 * w * h)
        return np.resize(data, (h, w, 2))


def _as_hw2(uv):
    if hasattr(uv, "detach"):                     # torch tensor (CPU or CUDA)
        uv = uv.detach().float().cpu().numpy()
    uv = np.asarray(uv)
    if uv.ndim == 3 and uv.shape[2] != 2 and uv.shape[0] == 2:     # the network's [2,H,W] -> [H,W,2]
        uv = uv.transpose(1, 2, 0)
    return uv


def writeFlow(filename, uv, v=None):
    """uv: [H,W,2] (or the network's [2,H,W]); or u and v given separately as [H,W] arrays."""
    if v is None:
        uv = _as_hw2(uv)
        assert uv.ndim == 3 and uv.shape[2] == 2
        u, v = uv[:, :, 0], uv[:, :, 1]
    else:
        u, v = np.asarray(_as_hw2(uv)), np.asarray(_as_hw2(v))
    assert u.shape == v.shape
    height, width = u.shape
    with open(filename, "wb") as f:
        f.write(TAG_CHAR.tobytes())
        np.array(width).astype(np.int32).tofile(f)
        np.array(height).astype(np.int32).tofile(f)
        np.stack((u, v), axis=2).astype(np.float32).tofile(f)


def flow_to_image_cuda(flow):
    """Middlebury colour coding on the GPU (ft_flow_to_image; flowlib.flow_to_image, lib/flownet/utils/flowlib.py:243-282).
    flow: CUDA fp32 [B,2,H,W] or [2,H,W] (the networks' output) -> CUDA uint8 [B,H,W,3] / [H,W,3] RGB."""
    import torch

    from ... import _lib, engine
    if not flow.is_cuda:
        raise RuntimeError("flow_to_image_cuda needs a CUDA tensor (no CPU path)")
    single = flow.dim() == 3
    f = (flow.unsqueeze(0) if single else flow).contiguous().float()
    B, two, H, W = f.shape
    assert two == 2
    with torch.cuda.device(f.device):
        img = torch.empty(B, H, W, 3, dtype=torch.uint8, device=f.device)
        ws = torch.empty(max(B, 1), dtype=torch.int32, device=f.device)
        _lib.check(_lib.lib().ft_flow_to_image(f.data_ptr(), B, H, W, ws.data_ptr(), img.data_ptr(), engine.stream_ptr()), "flow_to_image")
    return img[0] if single else img


def flow_to_image(flow):
    """flowlib.flow_to_image signature: flow ndarray [H,W,2] -> uint8 ndarray [H,W,3] (computed on the GPU)."""
    import torch
    f = torch.as_tensor(np.ascontiguousarray(np.asarray(flow, dtype=np.float32).transpose(2, 0, 1))).cuda()
    return flow_to_image_cuda(f).cpu().numpy()
