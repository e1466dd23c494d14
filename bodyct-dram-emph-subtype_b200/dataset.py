"""`SubtypingInference`: the reference's inference dataset (dataset.py:14-92) on the native .mha reader.

Per scan: read CT + lobe segmentation, lung = lobe > 0, blank everything outside the lung dilated by
two 3x3x3 steps (== one 5x5x5 box) to -2048, crop to the lung bounding box + 5 mm, derive the LAA-910
`ess` mask, and hand the dict to the transform.  With a CUDA `device` these pre-steps run on the GPU
(`ops.mask_bbox` + `ops.lung_crop`): the scan and the lobe labels are copied to the device
once (3 bytes per voxel), only the 6 bounding-box integers come back, and the cropped tensors go straight
into the GPU transform.  Without a device the same steps run in numpy/torch on the CPU.
"""
import glob
import math
import os
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

from . import mha_io
from .utils import find_crops


def grow_bbox(bbox, shape, spacing, border):
    """utils.py:55-60: pad the find_objects box by ceil(border/spacing) voxels per axis, clipped to the volume."""
    out = []
    for ax in range(3):
        lo, hi = int(bbox[2 * ax]), int(bbox[2 * ax + 1])
        pad = int(math.ceil(border / spacing[ax])) if border > 0 else 0
        out.append((max(0, lo - pad), min(int(shape[ax]), hi + pad)))
    return out


class SubtypingInference(torch.utils.data.Dataset):
    label_to_cle = {0: "absent", 1: "trace", 2: "mild", 3: "moderate", 4: "confluence", 5: "destructive"}
    label_to_pse = {0: "absent", 1: "mild", 2: "substantial"}

    def __init__(self, scan_path, lobe_path, transforms=None, keep_sorted=True, crop_border=5, device=None):
        super().__init__()
        self.device = torch.device(device) if device is not None else getattr(transforms, "device", None)
        self.scan_path, self.lobe_path = scan_path, lobe_path
        self.keep_sorted, self.transforms, self.crop_border = keep_sorted, transforms, crop_border
        self.scan_files = sorted(glob.glob(os.path.join(scan_path, "*.mha")))
        self.lobe_files = sorted(glob.glob(os.path.join(lobe_path, "*.mha")))
        self.scan_meta_cache = {}

    def __len__(self):
        return len(self.scan_files)

    def __getitem__(self, index):
        return self.get_data(index)

    def read_image(self, path):
        """(array [z,y,x], origin, spacing, direction) with the metadata reversed to z-y-x like dataset.py:49-55."""
        arr, meta = mha_io.read_mha(path)
        direction = np.asarray(meta["direction"]).reshape(3, 3)[::-1].flatten().tolist()
        return arr, meta["origin"][::-1], meta["spacing"][::-1], direction

    @staticmethod
    def dilate_lung(lung):
        """binary_dilation(lung, full 3x3x3 structure, iterations=2) == 5x5x5 box maximum, zero border."""
        t = torch.from_numpy(np.ascontiguousarray(lung)).to(torch.float32)[None, None]
        return F.max_pool3d(t, kernel_size=5, stride=1, padding=2)[0, 0].numpy() > 0

    def device_presteps(self, scan, lobe, spacing, uid):
        """dataset.py:66-83 on the GPU; returns the same dict with device tensors (masks as bool)."""
        from . import ops

        dev = self.device
        scan_d = torch.as_tensor(np.ascontiguousarray(scan)).to(torch.int16).to(dev, non_blocking=True)
        lobe_d = torch.as_tensor(np.ascontiguousarray(lobe))
        lobe_d = (lobe_d.to(torch.uint8) if lobe_d.dtype in (torch.uint8, torch.int8, torch.bool)
                  else (lobe_d > 0).to(torch.uint8)).to(dev, non_blocking=True)
        bbox = ops.mask_bbox(lobe_d).cpu().tolist()
        if bbox[1] <= bbox[0]:
            raise IndexError(f"{uid}: empty lobe segmentation (the reference fails in find_objects here)")
        crop = grow_bbox(bbox, scan.shape, spacing, self.crop_border)
        image, lung, ess = ops.lung_crop(scan_d, lobe_d, crop)
        return {
            "image": image,
            "lung_mask": lung.view(torch.bool),
            "ess_mask": ess.view(torch.bool),
            "crop_slice": np.asarray(crop),
            "original_size": np.asarray(scan.shape),
            "uid": uid,
        }

    def get_data(self, index):
        scan_file, lobe_file = self.scan_files[index], self.lobe_files[index]
        uid = Path(scan_file).stem
        scan, origin, spacing, direction = self.read_image(scan_file)
        lobe, _, _, _ = self.read_image(lobe_file)
        assert lobe.shape == scan.shape, "scan and lobe segmentation have different shapes."
        self.scan_meta_cache[uid] = {"spacing": spacing, "origin": origin, "direction": direction}
        if self.device is None and hasattr(self.transforms, "_dev") and torch.cuda.is_available():
            self.device = self.transforms._dev()  # the GPU transform follows: do the pre-steps there too
        if self.device is not None and self.device.type == "cuda":
            sample = self.device_presteps(scan, lobe, spacing, uid)
            return self.transforms(sample) if self.transforms else sample
        original = scan.copy()
        lung = lobe > 0
        scan = scan.copy()
        scan[~self.dilate_lung(lung)] = -2048
        sl = find_crops(lung, spacing, self.crop_border)
        scan_c, lung_c = scan[sl], lung[sl]
        sample = {
            "image": scan_c.astype(np.int16),
            "original_image": original[sl].astype(np.int16),
            "lung_mask": lung_c > 0,
            "ess_mask": np.logical_and(scan_c < -910, lung_c > 0),
            "crop_slice": np.asarray([(s.start, s.stop) for s in sl]),
            "original_size": np.asarray(scan.shape),
            "uid": uid,
        }
        self.scan_meta_cache[uid] = {"spacing": spacing, "origin": origin, "direction": direction}
        return self.transforms(sample) if self.transforms else sample
