"""re-export of the package's seeded synthetic-input generators (linesegmentdetector-slam_b200/synth.py)"""
import importlib.util
import os

_p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "linesegmentdetector-slam_b200", "synth.py")
_spec = importlib.util.spec_from_file_location("lsdb200_synth", _p)
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
occupancy_grid = _m.occupancy_grid
fake_scan_frame = _m.fake_scan_frame
lidar_frame = _m.lidar_frame
