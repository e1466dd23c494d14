"""Host-side helpers with the reference's names (utils.py of the reference):
`get_model_by_name` (utils.py:83-85), `load_state_dict_greedy` (226-249), `windowing` (28-37),
`find_crops` (53-63), `write_array_to_mha_itk` (87-104, here on the native MetaImage writer).
"""
import importlib
import logging
import math
import os

import numpy as np
import torch

from . import mha_io

logger = logging.getLogger(__name__)
PKG_DIR = os.path.dirname(os.path.abspath(__file__))


def _load_yaml(path):
    try:
        import yaml

        with open(path) as f:
            return yaml.safe_load(f)
    except ImportError:  # two-line configs: `_target_: a.b` and optionally `n_classes: [6, 3]`
        cfg = {}
        with open(path) as f:
            for line in f:
                if ":" in line:
                    k, v = line.split(":", 1)
                    v = v.strip()
                    cfg[k.strip()] = [int(t) for t in v.strip("[]").split(",")] if v.startswith("[") else v
        return cfg


def get_model_by_name(name):
    """conf/<name>.yaml -> module.  Looks in ./conf first (the reference resolves it relative to the
    working directory) and then next to this package; `_target_: med3d.<factory>` resolves to this
    package's med3d."""
    for base in (os.path.join(os.getcwd(), "conf"), os.path.join(PKG_DIR, "conf")):
        path = os.path.join(base, f"{name}.yaml")
        if os.path.isfile(path):
            break
    else:
        raise FileNotFoundError(f"no conf/{name}.yaml in {os.getcwd()} or {PKG_DIR}")
    cfg = dict(_load_yaml(path))
    target = cfg.pop("_target_")
    mod_name, fn_name = target.rsplit(".", 1)
    if mod_name == "med3d":
        from . import med3d as mod
    else:
        mod = importlib.import_module(mod_name)
    return getattr(mod, fn_name)(**cfg)


def load_state_dict_greedy(model, state_dict_to_load):
    """Copies every entry whose name and shape match, warns about the rest, never raises."""
    own = model.state_dict()
    for key, value in state_dict_to_load.items():
        if key not in own:
            logger.warning(f"[load_state_dict_greedy]:unexpected entry:{key}")
        elif own[key].shape != value.shape:
            logger.warning(f"[load_state_dict_greedy]:shape mismatch:{key}")
        else:
            own[key] = value
    for key in own:
        if key not in state_dict_to_load:
            logger.warning(f"[load_state_dict_greedy]:missing entry:{key}")
    model.load_state_dict(own, strict=False)


def windowing(image, from_span=(-1150, 350), to_span=(0, 255)):
    lo, hi = (np.min(image), np.max(image)) if from_span is None else from_span
    image = np.clip(image, a_min=lo, a_max=hi)
    return ((image - lo) / float(hi - lo)) * (to_span[1] - to_span[0]) + to_span[0]


def find_crops(mask, spacing, border):
    """Bounding box of mask>0 grown by ceil(border/spacing) voxels per axis and clipped to the volume."""
    nz = np.nonzero(np.asarray(mask) > 0)
    out = []
    for ax in range(len(nz)):
        lo, hi = int(nz[ax].min()), int(nz[ax].max()) + 1
        pad = int(math.ceil(border / spacing[ax])) if border > 0 else 0
        out.append(slice(max(0, lo - pad), min(mask.shape[ax], hi + pad)))
    return tuple(out)


def write_array_to_mha_itk(target_path, arrs, names, type=np.int16, origin=(0.0, 0.0, 0.0),
                           direction=tuple(np.eye(3, dtype=np.float64).flatten().tolist()),
                           spacing=(1.0, 1.0, 1.0), orientation="RAI"):
    """arr is z-y-x; origin/spacing/direction are given in ITK (x-y-z) order like the reference's call sites."""
    for arr, name in zip(arrs, names):
        mha_io.write_mha(os.path.join(target_path, f"{name}.mha"), np.asarray(arr).astype(type),
                         spacing=spacing, origin=origin, direction=direction, compress=True)


def expand_tensor_dims(t, expected_dim):
    while t.dim() < expected_dim:
        t = t.unsqueeze(0)
    return t


def squeeze_tensor_dims(t, expected_dim, squeeze_start_index=0):
    while t.dim() > expected_dim:
        t = t.squeeze(squeeze_start_index)
    return t


def cat_all_gather(tensors):
    import torch.distributed as dist

    gathered = [torch.ones_like(tensors) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, tensors, async_op=False)
    return torch.cat(gathered, dim=0)


def bind_to_gpu_cpus(gpu_index):
    """Host-side placement for the one-process-per-GPU runs: binds the calling thread (and the threads it starts later)
    to the CPUs NVML names as local to the GPU, so that pinned host buffers — first-touch allocated — and the copy
    threads sit on the GPU's NUMA node.  Returns a short description, or None when NVML / the cpuset does not allow it
    (nothing changes then).  DRAM_B200_NUMA=0 turns it off."""
    if os.environ.get("DRAM_B200_NUMA", "1").lower() in ("0", "off", "false", "no"):
        return None
    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:  # NVML counts physical devices; CUDA indices follow CUDA_VISIBLE_DEVICES
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            entry = ids[gpu_index]
            handle = pynvml.nvmlDeviceGetHandleByUUID(entry) if entry.startswith("GPU-") else \
                pynvml.nvmlDeviceGetHandleByIndex(int(entry))
        else:
            handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        after = sorted(os.sched_getaffinity(0))
        return f"nvml ideal CPUs of GPU {gpu_index}: {len(after)} of {before} ({after[0]}-{after[-1]})"
    except Exception as exc:  # no NVML, no permission, CPUs outside the container's cpuset ...
        logging.getLogger(__name__).debug("bind_to_gpu_cpus(%s) skipped: %s", gpu_index, exc)
        return None

