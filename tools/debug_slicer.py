import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hockey-vision-analytics_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib
t = importlib.import_module("test_gpu_team_e2e")
import inspect
src = inspect.getsource(t.test_slicer_device_path_matches_restated_inference_slicer)
src = src.replace("        assert len(got[f]) == len(rx) > 0\n", "        np.savez(os.path.join(ROOT, 'gpurun_out', 'slicer_dbg_%d.npz' % f), gx=got[f].xyxy, gc=got[f].confidence, rx=rx, rc=rc)\n        continue\n")
ns = dict(t.__dict__); ns["os"] = os; ns["ROOT"] = ROOT
exec(src, ns)
from hvb.runtime import get_context
try:
    ns["test_slicer_device_path_matches_restated_inference_slicer"](get_context(0))
except Exception as e:
    print("exc", repr(e)[:300])
