"""Writes tests/golden/model_namespace.json from the reference's own model_torch.py (run in the build container,
where /root/reference exists): state-dict keys/shapes, per-tensor sums of the default initialisation under
torch.manual_seed(0), and parameter counts, for the builders the scripts call."""
import importlib.util
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.test_dropin import BUILDS, FIXTURE, _summary  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_model_torch", "/root/reference/model_torch.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
with open(FIXTURE, "w") as f:
    json.dump({name: _summary(ref, name) for name in BUILDS}, f)
print("wrote", FIXTURE)
