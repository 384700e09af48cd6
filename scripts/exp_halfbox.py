"""Experiment: does the walk speed up when the particles (hence the tet records touched) are confined
to a fraction of the mesh?  Same mesh (c2, 128 MB of records), same particle count, workload box =
lower `frac` of the mesh in z.  Tests the effective-L2-capacity hypothesis behind a per-die split."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pumiumtally_b200.tally import PumiTally
from pumiumtally_b200.workload import CONFIGS, SyntheticWorkload

cfg = CONFIGS["c2"]; cells = cfg["cells"]; n = cfg["particles"]
dev = torch.device("cuda", 0)
for frac in (1.0, 0.75, 0.5, 0.25):
    box = (float(cells[0]), float(cells[1]), float(cells[2]) * frac)
    wl = SyntheticWorkload(box=box, num_particles=n, mean_length=cfg["mean_length"], backend="torch", device=dev)
    init = wl.initial_positions()
    eng = PumiTally.from_spec(f"box:{cells[0]},{cells[1]},{cells[2]}", n, device=0)
    s = torch.cuda.current_stream().cuda_stream
    eng.copy_initial_position_device(init.contiguous().data_ptr(), s)
    torch.cuda.synchronize()
    res = []
    for step in range(4):
        o, d, f, w = wl.next_step()
        st0 = eng.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.move_device(o.data_ptr(), d.data_ptr(), f.data_ptr(), w.data_ptr(), s)
        e1.record(); torch.cuda.synchronize()
        st1 = eng.stats()
        segs = st1["segments"] - st0["segments"]
        res.append((e0.elapsed_time(e1), segs))
    ms, segs = res[-1]
    print(json.dumps({"z_fraction": frac, "table_MB_touched": round(128 * frac), "ms": round(ms, 3), "segments": segs,
                      "ns_per_1k_segments": round(ms * 1e6 / segs * 1e3 / 1e3, 3), "gseg_s": round(segs / ms / 1e6, 2)}), flush=True)
    del eng
