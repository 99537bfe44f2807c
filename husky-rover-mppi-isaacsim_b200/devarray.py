"""Device-array plumbing (PyTorch is used for device memory and streams only).

`DeviceArray` stands in for the `wp.array` objects the reference controller exposes
(MPPI_isaac.py:445-487): callers use `.numpy()`, `.assign()`, `.zero_()` on them and hand them to other
GPU libraries through `__cuda_array_interface__` / DLPack.
"""
from __future__ import annotations

import numpy as np
import torch


class _RawCuda:
    """Minimal __cuda_array_interface__ carrier for memory owned by the C library."""

    def __init__(self, ptr: int, shape, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None,
        }


def view_device_memory(ptr: int, shape, device: torch.device, dtype=torch.float32) -> torch.Tensor:
    """Zero-copy torch view of library-owned device memory."""
    typestr = {torch.float32: "<f4", torch.int32: "<i4"}[dtype]
    with torch.cuda.device(device):
        return torch.as_tensor(_RawCuda(ptr, shape, typestr), device=device)


def to_device_f32(obj, device: torch.device) -> torch.Tensor:
    """Accepts a torch tensor, a NumPy array (H2D copy), or any object exporting
    __cuda_array_interface__ / DLPack (e.g. a Warp array: zero-copy, as `controller.Z_wp = dem_wp` in
    visual_terrain_stack_full_terrain.py:567).  Returns a flat contiguous float32 CUDA tensor."""
    if isinstance(obj, DeviceArray):
        t = obj.tensor
    elif isinstance(obj, torch.Tensor):
        t = obj
    elif isinstance(obj, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(obj, dtype=np.float32))
    elif hasattr(obj, "__cuda_array_interface__"):
        t = torch.as_tensor(obj, device=device)
    elif hasattr(obj, "__dlpack__"):
        t = torch.from_dlpack(obj)
    else:
        t = torch.as_tensor(np.asarray(obj, dtype=np.float32))
    t = t.to(device=device, dtype=torch.float32).contiguous().reshape(-1)
    return t


class DeviceArray:
    """A named float32 device buffer with the small slice of the wp.array API the reference callers use."""

    def __init__(self, tensor: torch.Tensor, on_read=None):
        self.tensor = tensor
        self._on_read = on_read          # hook that makes the content current (lazy outputs)

    # --- wp.array look-alikes
    def numpy(self) -> np.ndarray:
        if self._on_read is not None:
            self._on_read()
        return self.tensor.detach().cpu().numpy()

    def assign(self, src) -> None:
        src_t = to_device_f32(src, self.tensor.device)
        self.tensor.reshape(-1).copy_(src_t)

    def zero_(self) -> None:
        self.tensor.zero_()

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    @property
    def ptr(self) -> int:
        return self.tensor.data_ptr()

    @property
    def __cuda_array_interface__(self):
        if self._on_read is not None:
            self._on_read()
        return self.tensor.__cuda_array_interface__

    def __dlpack__(self, stream=None):
        return self.tensor.__dlpack__(stream=stream)

    def __dlpack_device__(self):
        return self.tensor.__dlpack_device__()

    def __len__(self):
        return self.tensor.shape[0]

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, device={self.tensor.device})"
