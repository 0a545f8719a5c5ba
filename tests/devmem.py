"""Device buffers for the device-resident tests.  On a GPU box: torch CUDA tensors (torch is the
device-memory plumbing).  Under the functional simulator (tests/emu, GDV_EMU=1) "device memory" is
host memory, so numpy arrays stand in and the same GDV_MEM_DEVICE code path of the product runs."""
from __future__ import annotations

import os

import numpy as np

EMU = os.environ.get("GDV_EMU") == "1"


class DevBuf:
    def __init__(self, shape, dtype, fill=None):
        self.np_dtype = np.dtype(dtype)
        if EMU:
            self.a = np.empty(shape, dtype=self.np_dtype)
            if fill is not None:
                self.a[...] = fill
        else:
            import torch
            tdt = getattr(torch, self.np_dtype.name)
            if fill is None:
                self.a = torch.empty(shape, dtype=tdt, device="cuda")
            else:
                self.a = torch.full(shape if isinstance(shape, tuple) else (shape,), fill, dtype=tdt, device="cuda")

    @property
    def ptr(self) -> int:
        return self.a.ctypes.data if EMU else self.a.data_ptr()

    def numpy(self) -> np.ndarray:
        return self.a.copy() if EMU else self.a.cpu().numpy()


def stream() -> int:
    if EMU:
        return 0
    import torch
    return torch.cuda.current_stream().cuda_stream


def synchronize() -> None:
    if not EMU:
        import torch
        torch.cuda.synchronize()
