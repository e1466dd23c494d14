"""CPU oracle of everything around the network on the hot path: the inference transforms, the
predict_step dRAM generation and the severity labels.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates, with plain torch CPU / numpy ops:
  intensity_window      functional.py:13-26 (args from models.py:60)
  standardize           intensity_transforms.py:104-114
  interpolate_image     spatial_transforms.py:55-75   (args from models.py:62)
  interpolate_mask      spatial_transforms.py:77-97
  inference_transform   models.py:55-63 + base.py:119-133 (dispatch by key substring)
  predict_step          models.py:430-450
  ratio_to_label        processor.py:34-38 with the maps of dataset.py:99-112
  lung_crop_sample      dataset.py:57-92 + utils.py:53-63 (the CPU pre-steps, "next" row f1)
  postprocess_scan      processor.py:111-143 + utils.py:28-37 ("next" row f2)
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import med3d_oracle as M

CLE_RATIO_MAP = {0: (0.0, 0.01), 1: (0.01, 0.05), 2: (0.05, 0.1), 3: (0.1, 0.2), 4: (0.2, 0.3), 5: (0.3, 1.0001)}
PSE_RATIO_MAP = {0: (0.0, 0.01), 1: (0.01, 0.05), 2: (0.05, 1.0001)}
DEFAULT_TARGET_SIZE = (128, 224, 288)  # processor.py:62


def intensity_window(img, from_span=(-1150, -300), to_span=(0, 1)):
    lo, hi = from_span
    x = torch.clamp(img.float(), min=lo, max=hi)
    return ((x - lo) / (hi - lo)) * (to_span[1] - to_span[0]) + to_span[0]


def standardize(x):
    x = x - x.mean()
    return x / x.std()  # torch.std: unbiased (N-1)


def slice_indices(d_in, d_out):
    return torch.linspace(0, d_in - 1, d_out).long()


def interpolate_image(x, target_size):
    """[D,H,W] fp32 -> target (D2,H2,W2): bilinear in-plane (align_corners=True), D by slice pick."""
    y = F.interpolate(x[None].float(), size=tuple(target_size[1:]), mode="bilinear", align_corners=True)
    return y[:, slice_indices(x.shape[0], target_size[0])][0].type(x.dtype)


def interpolate_mask(m, target_size):
    """[D,H,W] bool -> target: legacy `nearest` in-plane, D by the same slice pick, dtype restored."""
    y = F.interpolate(m[None].float(), size=tuple(target_size[1:]), mode="nearest")
    return y[:, slice_indices(m.shape[0], target_size[0])][0].type(m.dtype)


def inference_transform(sample, target_size=DEFAULT_TARGET_SIZE):
    """The TEST_PHASE Compose of models.py:55-63 applied to a dataset dict (numpy arrays in)."""
    out = {}
    for key, val in sample.items():
        t = torch.as_tensor(val) if isinstance(val, np.ndarray) else val  # NumpyToTensor
        if isinstance(t, torch.Tensor) and "image" in key:
            t = interpolate_image(standardize(intensity_window(t)), target_size)
        elif isinstance(t, torch.Tensor) and "mask" in key:
            t = interpolate_mask(t, target_size)
        out[key] = t
    return out


def predict_step(sd, arch, batch):
    """models.py:430-450 for a collated batch (image [B,D,H,W] fp32, masks [B,D,H,W] bool)."""
    scans = batch["image"].unsqueeze(1)
    lungs = batch["lung_mask"].unsqueeze(1).float()
    ess = batch["ess_mask"].unsqueeze(1).float()
    dense, _ = M.forward(sd, arch, scans, lungs)
    size = scans.shape[-3:]
    cle = F.interpolate(dense[0], size=size, mode="trilinear", align_corners=True) * ess
    pse = F.interpolate(dense[1], size=size, mode="trilinear", align_corners=True) * ess
    denom = lungs.sum()  # whole batch (quirk Q1)
    return {
        "cle_dense_outs": cle,
        "pse_dense_outs": pse,
        "cle_precentages": cle.view(cle.shape[0], -1).sum(-1) / denom,
        "pse_precentages": pse.view(pse.shape[0], -1).sum(-1) / denom,
        "crop_slices": batch.get("crop_slice"),
        "original_size": batch.get("original_size"),
        "uids": batch.get("uid"),
    }


def ratio_to_label(ratio, ratio_map):
    for label, (lo, hi) in ratio_map.items():
        if lo <= ratio < hi:
            return label
    raise IndexError(f"ratio {ratio} outside every bin")  # the reference raises IndexError too ([...][0])


def find_crops(mask, spacing, border):
    """utils.py:53-63: bounding box of mask>0, padded by ceil(border/spacing) voxels, clipped."""
    nz = np.nonzero(mask > 0)
    sl = []
    for ax in range(3):
        lo, hi = int(nz[ax].min()), int(nz[ax].max()) + 1
        pad = int(math.ceil(border / spacing[ax])) if border > 0 else 0
        sl.append(slice(max(0, lo - pad), min(mask.shape[ax], hi + pad)))
    return tuple(sl)


def lung_crop_sample(scan, lobe, spacing=(1.0, 1.0, 1.0), crop_border=5, uid="scan"):
    """dataset.py:57-92 without file I/O: numpy int16 scan + lobe labels -> the dataset dict."""
    from scipy import ndimage

    scan = np.array(scan, copy=True)
    original = scan.copy()
    lung = lobe > 0
    dlung = ndimage.binary_dilation(lung, ndimage.generate_binary_structure(3, 3), iterations=2)
    scan[~dlung] = -2048
    sl = find_crops(lung, spacing, crop_border)
    scan_c, lung_c = scan[sl], lung[sl]
    return {
        "image": scan_c.astype(np.int16),
        "original_image": original[sl].astype(np.int16),
        "lung_mask": lung_c > 0,
        "ess_mask": np.logical_and(scan_c < -910, lung_c > 0),
        "crop_slice": np.asarray([(s.start, s.stop) for s in sl]),
        "original_size": np.asarray(scan.shape),
        "uid": uid,
    }


def postprocess_scan(dense_out, crop_slice, original_size):
    """processor.py:115-122,143: resample one [1,D,H,W] dRAM to the crop, paste, window to uint8."""
    recon = tuple(int(s[1]) - int(s[0]) for s in crop_slice)
    up = F.interpolate(dense_out.unsqueeze(0), size=recon, mode="trilinear", align_corners=True)[0, 0].numpy()
    full = np.zeros(tuple(int(v) for v in original_size))
    full[tuple(slice(int(s[0]), int(s[1])) for s in crop_slice)] = up
    img = np.clip(full, 0, 1)
    img = ((img - 0) / float(1 - 0)) * 255 + 0  # utils.windowing(from_span=(0,1))
    return img.astype(np.uint8)
