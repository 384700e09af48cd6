"""End-to-end MoveToNextLocation from pinned host memory with and without delta upload of origins."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pumiumtally_b200.tally import PumiTally
from pumiumtally_b200.workload import CONFIGS, SyntheticWorkload

cfg = CONFIGS["c2"]; cells = cfg["cells"]; n = cfg["particles"]
box = tuple(float(c) for c in cells)
wl = SyntheticWorkload(box=box, num_particles=n, mean_length=cfg["mean_length"], backend="torch", device="cuda")
init = wl.initial_positions().cpu().numpy()
batches = [tuple(x.cpu().numpy() for x in wl.next_step()) for _ in range(6)]
ref = None
for mode, threads in ((0, 32), (2, 32), (2, 64), (2, 128)):
    eng = PumiTally.from_spec(f"box:{cells[0]},{cells[1]},{cells[2]}", n, device=0)
    eng.set_option("delta_upload", mode)
    eng.set_option("delta_threads", threads)
    eng.CopyInitialPosition(init.reshape(-1))
    bufs = [torch.empty(s, dtype=d, pin_memory=True) for s, d in ((3 * n, torch.float64), (3 * n, torch.float64), (n, torch.int8), (n, torch.float64))]
    O, D, F, W = (b.numpy() for b in bufs)
    times, host_us, saved = [], [], []
    s0 = eng.stats()["segments"]
    for o, d, f, w in batches:
        O[:], D[:], F[:], W[:] = o.reshape(-1), d.reshape(-1), f, w
        t0 = time.perf_counter()
        eng.MoveToNextLocation(O, D, F, W)
        segs = eng.stats()["segments"]
        times.append(time.perf_counter() - t0)
        host_us.append(eng.get_option("delta_host_us")); saved.append(eng.get_option("delta_saved_bytes"))
    ms = 1e3 * np.median(times[1:])
    flux = eng.flux
    if ref is None:
        ref = (flux, eng.elem_ids.copy(), eng.positions.copy())
        same = True
    else:
        same = bool(np.allclose(flux, ref[0], rtol=1e-9, atol=0) and (eng.elem_ids == ref[1]).all() and (eng.positions == ref[2]).all())
    print(json.dumps({"delta_upload": mode, "ms_per_move_median": round(ms, 2), "times_ms": [round(1e3 * t, 2) for t in times],
                      "host_compare_ms": [round(u / 1e3, 2) for u in host_us], "saved_MB": [round(s / 1e6, 1) for s in saved],
                      "h2d_MB_per_move": round(eng.stats()["h2d_bytes"] / len(batches) / 1e6, 1),
                      "gseg_s": round((segs - s0) / len(batches) / ms / 1e6, 2), "same_results_as_off": same,
                      "still_on": eng.get_option("delta_upload"), "threads": threads, "cpus": os.cpu_count()}), flush=True)
    del eng
