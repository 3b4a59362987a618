"""The reference's non_max_suppression (build_utils/utils.py:387-464, nc = 1, multi_label False) restated around
torchvision.ops.nms for GPU timing comparisons (tools/nms_sweep.py, tools/reference_gpu_fps.py)."""
import torch
import torchvision


def reference_gpu_nms(prediction, conf_thres, iou_thres, max_num=100):
    out = [None] * prediction.shape[0]
    for xi, x in enumerate(prediction):
        x = x[x[:, 4] > conf_thres]
        x = x[((x[:, 2:4] > 2) & (x[:, 2:4] < 4096)).all(1)]
        if not x.shape[0]:
            continue
        x = x.clone()
        x[..., 5:] *= x[..., 4:5]
        box = torch.stack((x[:, 0] - x[:, 2] / 2, x[:, 1] - x[:, 3] / 2, x[:, 0] + x[:, 2] / 2, x[:, 1] + x[:, 3] / 2), 1)
        conf, j = x[:, 5:].max(1)
        x = torch.cat((box, conf.unsqueeze(1), j.float().unsqueeze(1)), 1)[conf > conf_thres]
        if not x.shape[0]:
            continue
        boxes, scores = x[:, :4] + x[:, 5:6] * 4096, x[:, 4]
        i = torchvision.ops.nms(boxes, scores, iou_thres)[:max_num]
        out[xi] = x[i]
    return out
