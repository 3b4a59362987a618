"""Darknet ``.cfg`` / ``.data`` readers — the plugin format of the hot path.

Behavioural mirror of the reference's build_utils/parse_config.py:5-90 (same dialect, same value
typing, same errors) so that every shipped ``config/*.cfg`` yields the identical list of block dicts:

* blank lines and lines starting with ``#`` are dropped *before* stripping (an indented ``#`` line is
  therefore a syntax error in the reference too), ``;``-prefixed lines inside ``[net]`` become odd keys;
* ``anchors`` -> float ndarray (-1, 2); ``from`` / ``layers`` / ``mask`` (and ``size`` when it has a
  comma) -> list[int]; values for which ``str.isnumeric()`` holds -> int; everything else (``.7``,
  ``1.0``, ``leaky``) stays a string (parse_config.py:45-49);
* ``[convolutional]`` blocks start with ``batch_normalize = 0``;
* any key outside the supported list in a non-``[net]`` block raises ValueError (parse_config.py:52-63).
"""
import os

import numpy as np

SUPPORTED_KEYS = frozenset([
    'type', 'batch_normalize', 'filters', 'size', 'stride', 'pad', 'activation', 'layers', 'groups',
    'from', 'mask', 'anchors', 'classes', 'num', 'jitter', 'ignore_thresh', 'truth_thresh', 'random',
    'stride_x', 'stride_y', 'weights_type', 'weights_normalization', 'scale_x_y', 'beta_nms', 'nms_kind',
    'iou_loss', 'iou_normalizer', 'cls_normalizer', 'iou_thresh', 'probability', 'max_delta', 'atoms',
    'na', 'nc', 'squeeze_factor', 'n1x1', 'n3x3_reduce', 'n3x3', 'n5x5_reduce', 'n5x5', 'pool_proj'])

_INT_LIST_KEYS = ('from', 'layers', 'mask')


def _typed(key: str, val: str):
    if key == 'anchors':
        return np.array([float(v) for v in val.replace(' ', '').split(',')]).reshape((-1, 2))
    if key in _INT_LIST_KEYS or (key == 'size' and ',' in val):
        return [int(v) for v in val.split(',')]
    if val.isnumeric():  # digits only: floats such as ".7" or "1.0" deliberately stay strings
        return int(val)
    return val


def parse_model_cfg_text(text: str) -> list:
    """Parses cfg text (see parse_model_cfg)."""
    blocks = []
    for raw in text.split('\n'):
        if not raw or raw.startswith('#'):
            continue
        line = raw.strip()
        if line.startswith('['):
            block = {'type': line[1:-1].strip()}
            if block['type'] == 'convolutional':
                block['batch_normalize'] = 0
            blocks.append(block)
            continue
        key, val = line.split('=')  # ValueError on malformed lines, as in the reference
        key, val = key.strip(), val.strip()
        blocks[-1][key] = _typed(key, val)
    for block in blocks[1:]:
        for key in block:
            if key not in SUPPORTED_KEYS:
                raise ValueError("Unsupported fields:{} in cfg".format(key))
    return blocks


def parse_model_cfg(path: str) -> list:
    """``.cfg`` file -> [net-dict, block-dict, ...]   (reference parse_config.py:5-65)."""
    if not path.endswith('.cfg') or not os.path.exists(path):
        raise FileNotFoundError("the cfg file not exist...")
    with open(path, 'r', encoding='utf-8') as f:
        return parse_model_cfg_text(f.read())


def parse_data_cfg(path: str) -> dict:
    """``.data`` file of ``key = value`` lines -> dict of strings (reference parse_config.py:68-90)."""
    if not os.path.exists(path) and os.path.exists('data' + os.sep + path):
        path = 'data' + os.sep + path
    options = {}
    with open(path, 'r') as f:
        for line in f:
            line = line.strip()
            if line == '' or line.startswith('#'):
                continue
            key, val = line.split('=')
            options[key.strip()] = val.strip()
    return options
