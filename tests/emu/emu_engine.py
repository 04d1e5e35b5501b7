"""CaptureEngine over the CPU emulation build of the kernel sources -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).

The product engine (mocapv2_b200/engine.py) refuses to exist without a CUDA device and has no CPU branch; this subclass
swaps the device plumbing (streams, pinned staging, device context) for host equivalents so that the GPU-less part of the
test-suite can drive the SAME host logic and the SAME kernel sources (compiled by build_emu.py) against the oracle.
"""
import contextlib
import ctypes

import torch

from mocapv2_b200 import _cabi
from mocapv2_b200.engine import CaptureEngine


class EmuEngine(CaptureEngine):
    def __init__(self, lib_path):
        self.device = torch.device("cpu")
        self.lib = _cabi.load(lib_path)
        self._init_state()

    def _stream(self):
        return ctypes.c_void_p(0)

    def _device_ctx(self):
        return contextlib.nullcontext()

    def _retire(self, t):
        pass

    def thread_stream(self):
        return contextlib.nullcontext()

    def upload_image(self, a):
        return torch.from_numpy(a)[None].clone()

    def download_image_async(self, t):
        out = t.clone().numpy()
        return lambda: out.copy()
