"""ORACLE — TEST INFRASTRUCTURE ONLY.  numpy fp32 restatement of the reference's post-NMS box rescaling:
build_utils/utils.py:60-84 scale_coords and :87-92 clip_coords (in-place tensor ops: subtract the padding, divide by the
gain, clamp to the original image).  Pinned by tests/golden/coords_cases.npz (tests/golden/make_coords_golden.py runs the
reference's own functions)."""
import numpy as np


def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None):
    c = np.array(coords, dtype=np.float32, copy=True)
    if ratio_pad is None:
        gain = max(img1_shape) / max(img0_shape)
        pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    else:
        gain, pad = ratio_pad[0][0], ratio_pad[1]
    c[:, [0, 2]] -= np.float32(pad[0])
    c[:, [1, 3]] -= np.float32(pad[1])
    c[:, :4] /= np.float32(gain)
    return clip_coords(c, img0_shape)


def clip_coords(boxes, img_shape):
    b = np.array(boxes, dtype=np.float32, copy=True)
    b[:, 0] = np.clip(b[:, 0], np.float32(0), np.float32(img_shape[1]))
    b[:, 1] = np.clip(b[:, 1], np.float32(0), np.float32(img_shape[0]))
    b[:, 2] = np.clip(b[:, 2], np.float32(0), np.float32(img_shape[1]))
    b[:, 3] = np.clip(b[:, 3], np.float32(0), np.float32(img_shape[0]))
    return b
