#!/usr/bin/env python
"""Regenerates the reference's missing `monteCarlo/mergedGridSearchResultFinal.npy` on the GPU
 and writes it next to the path the reference config names.

    python tools/regenerate_table.py [out.npy] [--precision fp64|fp32]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import mdr_b200
    out = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "mergedGridSearchResultFinal.npy"
    precision = "fp32" if "--precision" in sys.argv and sys.argv[sys.argv.index("--precision") + 1] == "fp32" else "fp64"
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    table = mdr_b200.regenerate_table(precision=precision)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    np.save(out, table)
    n = table.size * mdr_b200.montecarlo.NB_TIME_STEPS_BY_SIM
    print("table %s: %d entries, %.2f s (%.3g house-steps/s incl. host population building), mean %.1f W, "
          "zeros %.1f%%, saved to %s" % (precision, table.size, dt, n / dt, table.mean(), 100 * (table == 0).mean(), out))


if __name__ == "__main__":
    main()
