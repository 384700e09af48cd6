"""Cost of the score filter: device-resident c2 moves with 1, 2, 4, 8 score bins (one masked pass of the
walk kernel per bin + one for the unscored particles), and one host-pointer binned move.
Usage: python scripts/exp_score_bins.py [config] [particles]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pumiumtally_b200.tally import PumiTally
from pumiumtally_b200.workload import CONFIGS, SyntheticWorkload

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = CONFIGS[cfg_name]; cells = cfg["cells"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["particles"]
box = tuple(float(c) for c in cells)
dev = torch.device("cuda", 0)
steps, warm = 6, 3
stream = torch.cuda.current_stream().cuda_stream
for nbins in (1, 2, 4, 8):
    wl = SyntheticWorkload(box=box, num_particles=n, mean_length=cfg["mean_length"], mu_min=cfg["mu_min"], backend="torch", device=dev)
    eng = PumiTally.from_spec(f"box:{cells[0]},{cells[1]},{cells[2]}", n, device=0)
    eng.set_option("autotune", 0)
    eng.set_score_bins(nbins)
    eng.copy_initial_position_device(wl.initial_positions().contiguous().data_ptr(), stream)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms, segs = [], []
    for k in range(warm + steps):
        o, d, f, w = (x.contiguous() for x in wl.next_step())
        bins = torch.randint(0, nbins, (n,), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        s0 = eng.stats()
        ev0.record()
        eng.move_device_binned(o.data_ptr(), d.data_ptr(), f.data_ptr(), w.data_ptr(), bins.data_ptr(), stream)
        ev1.record()
        torch.cuda.synchronize()
        if k >= warm:
            ms.append(ev0.elapsed_time(ev1)); segs.append(eng.stats()["segments"] - s0["segments"])
    t = float(np.median(ms))
    print(json.dumps({"config": cfg_name, "particles": n, "score_bins": nbins, "ms_per_move": round(t, 3),
                      "gseg_s": round(float(np.median(segs)) / t / 1e6, 2), "launches_per_move": "mask + walk per bin (+1 unscored)" if nbins > 1 else "walk"}), flush=True)
    if nbins == 4:  # the same through the host-pointer call (pageable arrays, bins uploaded ahead of the particle data)
        wlh = SyntheticWorkload(box=box, num_particles=n, mean_length=cfg["mean_length"], mu_min=cfg["mu_min"])
        e2 = PumiTally.from_spec(f"box:{cells[0]},{cells[1]},{cells[2]}", n, device=0)
        e2.set_score_bins(nbins)
        e2.CopyInitialPosition(wlh.initial_positions().reshape(-1))
        O, D, F, W = np.empty(3 * n), np.empty(3 * n), np.empty(n, dtype=np.int8), np.empty(n)
        B = np.random.default_rng(1).integers(0, nbins, n, dtype=np.int32)
        tt = {"binned": [], "plain": []}
        for k in range(8):
            o, d, f, w = wlh.next_step()
            O[:], D[:], F[:], W[:] = o.reshape(-1), d.reshape(-1), f, w
            e2.synchronize()
            t0 = time.perf_counter()
            if k % 2 == 0: e2.MoveToNextLocationBinned(O, D, F, W, B)
            else: e2.MoveToNextLocation(O, D, F, W)
            e2.stats()
            if k >= 2: tt["binned" if k % 2 == 0 else "plain"].append(1e3 * (time.perf_counter() - t0))
        print(json.dumps({"host_pointer_move_ms": {k_: round(float(np.median(v)), 2) for k_, v in tt.items()}, "score_bins": nbins}), flush=True)
        del e2
    del eng
