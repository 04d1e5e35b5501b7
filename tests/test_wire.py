"""Wire format of the realtime loop (RealtimeTracking_FLIR.py:184-191): byte-identical to msgpack.packb of the reference's dict."""
import numpy as np
import pytest
import torch

from mocapv2_b200 import synth as S
from mocapv2_b200.wire import TrackerPacketizer, pack_tracker
from util import lists_to_arrays

msgpack = pytest.importorskip("msgpack")


def test_packet_bytes_equal_msgpack():
    for point in ([0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0.2102, 0.4108, -2.1], [0, 0, 0, 0, 1e-300, -0.0, 3.5e12]):
        assert pack_tracker(point) == msgpack.packb({"tracker1": point}, use_bin_type=True)


def test_packetizer_follows_the_tracker_loop(engine):
    rig = S.config_rig("c1")
    cams = engine.cameras(rig["poses"], rig["camera_params"])
    Fs = torch.from_numpy(np.array(rig["Fs"], dtype=np.float64)).to(engine.device)
    pk = TrackerPacketizer(engine.device)
    xy, cnt = lists_to_arrays([[[None, None]], [[None, None]]])
    empty = engine.correspond(torch.from_numpy(xy).to(engine.device), torch.from_numpy(cnt).to(engine.device), Fs, cams, obj_count=4, fp64=True)
    assert pk.packet(empty) == msgpack.packb({"tracker1": [0, 0, 0, 0, 0, 0, 0, 0]}, use_bin_type=True)
    xy, cnt = lists_to_arrays([[[201, 184], [346, 97]], [[149, 366], [332, 247]]])
    res = engine.correspond(torch.from_numpy(xy).to(engine.device), torch.from_numpy(cnt).to(engine.device), Fs, cams, obj_count=4, fp64=True)
    assert int(res.n_obj[0]) > 0
    want = [0, 0, 0, 0] + list(res.obj[0, 0].cpu().numpy())
    assert pk.packet(res) == msgpack.packb({"tracker1": [float(v) if i >= 4 else v for i, v in enumerate(want)]}, use_bin_type=True)
    assert pk.packet(empty) == pk.packet(res)             # nothing new: the previous point is sent again
