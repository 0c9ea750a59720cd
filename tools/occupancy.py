"""Print how many clusters of the persistent recurrence kernels are co-resident on this GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctypes import byref, c_int
import torch
from biear_b200 import _lib
torch.cuda.init(); torch.zeros(1, device="cuda")
lib = _lib.load()
f, b = c_int(0), c_int(0)
_lib.check(lib.biear_adaptive_occupancy(100, 513, byref(f), byref(b)), "occupancy")
print("max active clusters: fwd", f.value, "bwd", b.value, "SMs", torch.cuda.get_device_properties(0).multi_processor_count)
