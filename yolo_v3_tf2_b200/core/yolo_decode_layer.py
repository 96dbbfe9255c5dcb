"""Drop-in for reference core/yolo_decode_layer.py: one CUDA kernel instead of ~25 TF ops."""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def _as_device_f32(t, device):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32))
    if not t.is_cuda:
        t = t.to(device, non_blocking=True)
    return t.contiguous().float()


def yolo_decode(model_output_grids, anchors_table, nclasses, with_scores=False, compact=False):
    """reference core/yolo_decode_layer.py:15-36.

    model_output_grids: list of [B, gh, gw, 3, 5+nclasses] float32 (torch CUDA tensors; numpy is copied to the GPU)
    anchors_table:      [n_scales, 3, 2] image-fraction anchors (reference core/utils.py:31-37)
    returns (bboxes [B,N,4], confidence [B,N,1], class_probs [B,N,nclasses]) as CUDA tensors; with
    ``with_scores=True`` additionally (scores [B,N], class_indices [B,N] int64) computed in the same pass
    (reference core/yolo_nms.py:18-24).  ``compact=True`` (implies with_scores; used by ``Detector``): confidence and
    class_probs are not written at all -- the return value is (bboxes, None, None, scores, class_indices) -- because
    yolo_nms only consumes boxes, scores and class ids.
    """
    ctx = _lib.context()
    dev = torch.device("cuda", ctx.device)
    grids = [_as_device_f32(g, dev) for g in model_output_grids]
    ns = len(grids)
    if ns < 1 or ns > 3:
        raise _lib.Y3Unsupported("1..3 output grids supported")
    F = 5 + int(nclasses)
    # 4-D grids [B, gh, gw, P >= 3*F] are the pitched layout of ``model(x, padded=True)``
    pitched = all(g.dim() == 4 for g in grids)
    for g in grids:
        if pitched:
            if g.shape[3] < 3 * F or g.shape[3] % 4 != 0:
                raise ValueError(f"pitched grid shape {tuple(g.shape)}: pitch must be >= {3 * F} and a multiple of 4")
        elif g.dim() != 5 or g.shape[3] != 3 or g.shape[4] != F:
            raise ValueError(f"grid shape {tuple(g.shape)} is not [B, gh, gw, 3, {F}]")
    B = grids[0].shape[0]
    anchors = np.ascontiguousarray(np.asarray(
        anchors_table.cpu().numpy() if isinstance(anchors_table, torch.Tensor) else anchors_table, dtype=np.float32))
    if anchors.shape != (ns, 3, 2):
        raise ValueError(f"anchors_table shape {anchors.shape} != ({ns}, 3, 2)")
    N = sum(int(g.shape[1] * g.shape[2] * 3) for g in grids)
    bboxes = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
    conf = probs = None
    if compact:
        with_scores = True
    else:
        conf = torch.empty((B, N, 1), dtype=torch.float32, device=dev)
        probs = torch.empty((B, N, int(nclasses)), dtype=torch.float32, device=dev)
    scores = cls = None
    if with_scores:
        scores = torch.empty((B, N), dtype=torch.float32, device=dev)
        cls = torch.empty((B, N), dtype=torch.int64, device=dev)
    gp = (C.c_void_p * ns)(*[g.data_ptr() for g in grids])
    gh = (C.c_int * ns)(*[int(g.shape[1]) for g in grids])
    gw = (C.c_int * ns)(*[int(g.shape[2]) for g in grids])
    if pitched:
        pp = (C.c_int * ns)(*[int(g.shape[3]) for g in grids])
        _lib.check(_lib.lib().y3_decode_pitched(ctx.handle, gp, gh, gw, pp, ns, anchors.ctypes.data_as(C.c_void_p), B,
                                                int(nclasses), _lib.ptr(bboxes), _lib.ptr(conf), _lib.ptr(probs),
                                                _lib.ptr(scores), _lib.ptr(cls), _lib.stream_ptr()))
    else:
        _lib.check(_lib.lib().y3_decode(ctx.handle, gp, gh, gw, ns, anchors.ctypes.data_as(C.c_void_p), B, int(nclasses),
                                        _lib.ptr(bboxes), _lib.ptr(conf), _lib.ptr(probs), _lib.ptr(scores), _lib.ptr(cls),
                                        _lib.stream_ptr()))
    if with_scores:
        return bboxes, conf, probs, scores, cls
    return bboxes, conf, probs
