"""Evaluation and on-disk artefacts of the label-estimation scripts:
the 2-class confusion / IoU the reference computes with chainercv
(batch_spalign_kmeans.py:398-405) on the GPU, and the files ``save_info`` writes
(:392-396 ``<name>.npy`` uint8 road mask, ``<name>_all_cluster.npy`` uint8 cluster map,
:407-422 one JSON line per image in ``result.json``) so that ``utils/mean_result.py`` and the
SegNet trainer downstream read them unchanged."""
from __future__ import annotations

import json
import os
import time
from typing import Optional

import numpy as np
import torch

from . import ops


def create_label_mask(label):
    """Cityscapes labelIds -> {-1 void, 1 road, 0 other} (batch_spalign_kmeans.py:279-296)."""
    label = np.asarray(label)
    ids = np.zeros(label.shape, dtype=np.int32)
    ids[label <= 6] = -1
    ids[label == 7] = 1
    return ids


def road_scores(road_masks, labels):
    """Per-image dicts with the reference's result.json score fields, computed from the 2x2
    confusion (rows = ground truth, columns = prediction, pixels with label < 0 ignored).

    road_masks [n, H, W] bool/uint8, labels [n, H, W] int (-1/0/1); NumPy or torch CUDA."""
    dev = road_masks.device if isinstance(road_masks, torch.Tensor) else torch.device('cuda', 0)
    rm = road_masks if isinstance(road_masks, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(road_masks))
    gt = labels if isinstance(labels, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(labels))
    conf = ops.confusion2(rm.to(dev), gt.to(dev)).cpu().numpy()
    out = []
    for c in conf:
        tp, fp, fn = int(c[1, 1]), int(c[0, 1]), int(c[1, 0])
        with np.errstate(invalid='ignore', divide='ignore'):
            iou = np.diag(c) / (c.sum(axis=1) + c.sum(axis=0) - np.diag(c)).astype(np.float64)
        out.append({
            'road_iou': float(iou[1]), 'non_road_iou': float(iou[0]),
            'precision': float(tp / (tp + fp)) if tp + fp > 0 else None,
            'recall': float(tp / (tp + fn)) if tp + fn > 0 else None,
            'TP': tp, 'FP': fp, 'FN': fn})
    return out


def save_info(out_dir, img_fn, label_fn, road_mask, clustering_result, scores, args=None,
              elapsed_times: Optional[dict] = None, st_all: Optional[float] = None):
    """Write the three artefacts of the reference's save_info for one image and return the
    JSON record.  ``scores`` is one entry of ``road_scores``."""
    os.makedirs(out_dir, exist_ok=True)
    stem = os.path.splitext(os.path.basename(img_fn))[0]
    rm = road_mask.cpu().numpy() if isinstance(road_mask, torch.Tensor) else np.asarray(road_mask)
    cr = clustering_result.cpu().numpy() if isinstance(clustering_result, torch.Tensor) \
        else np.asarray(clustering_result)
    np.save(os.path.join(out_dir, stem), rm.astype(np.uint8))
    np.save(os.path.join(out_dir, stem + '_all_cluster'), cr.astype(np.uint8))
    info = {'img_fn': img_fn, 'label_fn': label_fn}
    info.update(scores)
    if args is not None:
        info.update({k: v for k, v in vars(args).items()})
    times = dict(elapsed_times or {})
    if st_all is not None:
        times['elapsed_time'] = time.time() - st_all
    info.update(times)
    with open(os.path.join(out_dir, 'result.json'), 'a') as fp:
        print(json.dumps(info), file=fp)
    return info


# ---------------------------------------------------------------------------------------------
# Archives the downstream SegNet trainer reads
# ---------------------------------------------------------------------------------------------
def zip_estimated_labels(result_dir, zip_fn, pattern='*leftImg8bit.npy', root=None):
    """``find <result_dir> -name "*leftImg8bit.npy" | zip -0r <zip_fn> -@`` (README.md:134-136):
    an UNCOMPRESSED zip whose members keep the path ``find`` prints (relative to ``root``, default
    the current directory).  datasets/zipped_estimated_cityscapes_dataset.py:20-24,65-66 opens it
    with ``np.load`` and indexes members by that path.  Returns the member names."""
    import fnmatch
    import zipfile
    root = os.getcwd() if root is None else root
    names = []
    for d, _, files in sorted(os.walk(result_dir)):
        for fn in sorted(files):
            if fnmatch.fnmatch(fn, pattern):
                names.append(os.path.join(d, fn))
    with zipfile.ZipFile(zip_fn, 'w', compression=zipfile.ZIP_STORED, allowZip64=True) as zf:
        out = []
        for path in names:
            arc = os.path.relpath(path, root)
            zf.write(path, arcname=arc)
            out.append(arc)
    return out


class LabelZipWriter:
    """The same archive written directly from masks (no intermediate .npy files): one stored
    ``<prefix>/<stem>.npy`` member per image, uint8 like save_info (:393-394)."""

    def __init__(self, zip_fn, prefix='results/estimated_train_labels'):
        import zipfile
        self.zf = zipfile.ZipFile(zip_fn, 'w', compression=zipfile.ZIP_STORED, allowZip64=True)
        self.prefix = prefix
        self.names = []

    def add(self, img_fn, road_mask, suffix=''):
        import io
        stem = os.path.splitext(os.path.basename(img_fn))[0] + suffix
        rm = road_mask.cpu().numpy() if isinstance(road_mask, torch.Tensor) else np.asarray(road_mask)
        buf = io.BytesIO()
        np.save(buf, rm.astype(np.uint8))
        arc = '%s/%s.npy' % (self.prefix, stem)
        self.zf.writestr(arc, buf.getvalue())
        self.names.append(arc)
        return arc

    def close(self):
        self.zf.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def save_label_npz(zip_fn, preds_and_scores):
    """``np.savez(fp, **d)`` of utils/run_train_rounds.py:191-203: one archive holding, per
    image, the predicted label and its score map under the keys the relabel workers queue."""
    with open(zip_fn, 'wb') as fp:
        np.savez(fp, **{k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                        for k, v in preds_and_scores.items()})
    return zip_fn
