"""Worker process of the CPU baseline's CC pool: runs the UNMODIFIED reference's compute_cross_correlation_feature
(utils.py:390-420, from the oracle/_ref staging) on clips streamed over stdin / stdout.  Measurement infrastructure only.

Protocol (binary, little endian): request = int64 n, float64 fs, int64 num_lags, float64 max_lag_ms, n float32 left,
n float32 right; response = num_lags float32.  n == 0 ends the worker.  The reference runs this function in a
ProcessPoolExecutor over files (create_h5_data/data_save.py:213-221); plain pipes are used here because a
multiprocessing pool re-imports the parent's main module (torch) in every worker."""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from oracle import stage_ref
    _, ref_utils = stage_ref.load_reference(model=False)
    inp, out = sys.stdin.buffer, sys.stdout.buffer
    out.write(b"R")          # ready
    out.flush()
    while True:
        head = inp.read(32)
        if len(head) < 32:
            return
        n, fs, num_lags, max_lag_ms = struct.unpack("<qdqd", head)
        if n == 0:
            return
        left = np.frombuffer(inp.read(4 * n), np.float32)
        right = np.frombuffer(inp.read(4 * n), np.float32)
        cc = ref_utils.compute_cross_correlation_feature(left, right, fs, int(num_lags), max_lag_ms)
        out.write(np.asarray(cc, np.float32).tobytes())
        out.flush()


if __name__ == "__main__":
    main()
