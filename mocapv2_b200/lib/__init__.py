"""Drop-in mirror of the reference's `lib` package for the capture hot path: same module names
(`lib.ImageOperations`, `lib.Helpers`), same function names, arguments, return conventions and module globals, so
`from mocapv2_b200.lib.ImageOperations import _find_dot` replaces `from lib.ImageOperations import _find_dot`
(RealtimeTracking_FLIR.py:6-7, CapturePoints.py:5-6, CalculateCameraPoses.py:3)."""
