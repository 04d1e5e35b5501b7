"""The decomposition the next border stage rests on: every border step of cv2.findContours follows
from a 2x2 corner configuration, and per-(foreground, background)-label sums over the steps give the oracle's contour
records.  CPU only; checks tests/links_probe.py (a design probe, not product code) against cv2 and oracle/restate.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def test_border_steps_are_local_and_sum_to_the_contour_records(monkeypatch):
    import links_probe
    monkeypatch.setattr(sys, "argv", ["links_probe.py", "25"])
    assert links_probe.main() == 0
