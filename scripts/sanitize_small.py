"""Small end-to-end run of every shipped kernel variant for compute-sanitizer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import box_case, run_workload
from oracle.oracle import OraclePumiTally
from pumiumtally_b200.tally import PumiTally

variants = [int(v) for v in sys.argv[1:]] or [8, 16, 24, 0]
for v in variants:
    coords, t2v, wl = box_case((6, 6, 5), 6000)
    e = PumiTally.from_arrays(coords, t2v, wl.n)
    e.set_option("variant", v)
    e.set_option("chunk", 2048)  # several upload/compute ranges, ragged last chunk
    run_workload(e, OraclePumiTally(coords, t2v, wl.n), wl, steps=2, label=f"sanitize v{v}")
    print("variant", v, "ok", e.stats()["segments"], flush=True)
