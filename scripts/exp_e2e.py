"""End-to-end MoveToNextLocation from host memory: pageable vs cudaHostRegister'ed vs pinned buffers."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pumiumtally_b200.tally import PumiTally
from pumiumtally_b200.workload import CONFIGS, SyntheticWorkload

cfg = CONFIGS["c2"]; cells = cfg["cells"]; n = cfg["particles"]
box = tuple(float(c) for c in cells)
wl = SyntheticWorkload(box=box, num_particles=n, mean_length=cfg["mean_length"], backend="torch", device="cuda")
init = wl.initial_positions().cpu().numpy()
batches = [tuple(x.cpu().numpy() for x in wl.next_step()) for _ in range(5)]
for mode in ("pageable", "registered", "pinned"):
    eng = PumiTally.from_spec(f"box:{cells[0]},{cells[1]},{cells[2]}", n, device=0)
    eng.set_option("register_host", 1 if mode == "registered" else 0)
    eng.CopyInitialPosition(init.reshape(-1))
    if mode == "pinned":
        bufs = [torch.empty(s, dtype=d, pin_memory=True) for s, d in ((3 * n, torch.float64), (3 * n, torch.float64), (n, torch.int8), (n, torch.float64))]
        O, D, F, W = (b.numpy() for b in bufs)
    else:
        O, D, F, W = np.empty(3 * n), np.empty(3 * n), np.empty(n, dtype=np.int8), np.empty(n)
    times = []
    s0 = eng.stats()["segments"]
    for o, d, f, w in batches:
        O[:], D[:], F[:], W[:] = o.reshape(-1), d.reshape(-1), f, w
        t0 = time.perf_counter()
        eng.MoveToNextLocation(O, D, F, W)
        segs = eng.stats()["segments"]
        times.append(time.perf_counter() - t0)
    ms = 1e3 * np.median(times[1:])
    print(json.dumps({"mode": mode, "ms_per_move_median": round(ms, 2), "first_move_ms": round(1e3 * times[0], 1),
                      "GB_per_s": round(n * 57 / ms / 1e6, 1), "gseg_s": round((segs - s0) / len(batches) / ms / 1e6, 2)}), flush=True)
    del eng
