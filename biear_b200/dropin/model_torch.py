"""`import model_torch` for the reference's scripts: resolves to the B200 implementation.

Put this directory first on sys.path; train_biear.py:12 / evaluate_biear.py:11 then get
biear_b200's builders, classes and constants under the names they import."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(1, _ROOT)

from biear_b200.model_torch import *  # noqa: F401,F403,E402
from biear_b200.model_torch import (DATA_DIM, LATENT_DIM, N_DIST_CLASS, N_SECTORS, build_model,  # noqa: F401,E402
                                    build_model_active, build_model_active_single_controller,
                                    build_model_auralnet_active)

# BIEAR_TIMING=1: report at exit how much device / host time the front-end took per call (the reference's scripts have no
# timers of their own; used by tools/run_reference_pipeline.py for the unchanged-script runs)
if os.environ.get("BIEAR_TIMING") == "1":
    import atexit

    from biear_b200 import timing as _timing
    _timing.enabled = True
    atexit.register(_timing.report)
