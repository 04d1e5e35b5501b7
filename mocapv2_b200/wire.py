"""Output wire format of the realtime loop: the msgpack packet the reference's tracker thread sends to
Unity over TCP 127.0.0.1:5000 -- `msgpack.packb({"tracker1": [0, 0, 0, 0, x, y, z]}, use_bin_type=True)`
(RealtimeTracking_FLIR.py:184-191; same schema in unity_communication/ToUnity.py:15-17,85).

Host-side plumbing around the GPU path: `TrackerPacketizer` reads the top-ranked object point of a frame-set back with an
asynchronous device-to-host copy into pinned memory and formats the identical packet.  The socket itself (camera SDK,
Unity connection handling) stays with the caller, as in the reference.
"""
from __future__ import annotations

import struct

import torch


def pack_tracker(point) -> bytes:
    """msgpack bytes of {"tracker1": point} exactly as msgpack.packb(..., use_bin_type=True) produces them: a fixmap of one
    entry, the fixstr key, a fixarray whose ints are positive fixints and whose floats are float64.  `point` is the
    reference's list: 8 zeros before the first detection (RealtimeTracking_FLIR.py:171), then [0, 0, 0, 0, x, y, z]."""
    out = bytearray(b"\x81\xa8tracker1")
    if len(point) > 15:
        raise ValueError("tracker point lists are short (fixarray)")
    out.append(0x90 | len(point))
    for v in point:
        if isinstance(v, int) and not isinstance(v, bool):
            if not 0 <= v <= 127:
                raise ValueError("only the reference's zero padding is expected as an integer")
            out.append(v)
        else:
            out += b"\xcb" + struct.pack(">d", float(v))
    return bytes(out)


class TrackerPacketizer:
    """Per frame-set: best object point (device) -> pinned host buffer (async) -> packet bytes."""

    def __init__(self, device):
        self.device = torch.device(device)
        pin = self.device.type == "cuda"
        self._host = torch.zeros(4, dtype=torch.float64, pin_memory=pin)
        self._dev = torch.zeros(4, dtype=torch.float64, device=self.device)
        self.point = [0, 0, 0, 0, 0, 0, 0, 0]                 # RealtimeTracking_FLIR.py:171

    def packet(self, corr, s: int = 0) -> bytes:
        """corr: CorrespondResult of engine.correspond / pipeline.step; s: frame-set index.  Keeps the previous point when
        nothing was triangulated, like the reference's tracker loop (:185-189)."""
        self._dev[:3] = corr.obj[s, 0]
        self._dev[3] = corr.n_obj[s].to(torch.float64)
        self._host.copy_(self._dev, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        if self._host[3] > 0:
            self.point = [0, 0, 0, 0] + [float(v) for v in self._host[:3]]
        return pack_tracker(self.point)
