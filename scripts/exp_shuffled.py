"""Does the internal spatial renumbering make an arbitrarily numbered mesh as fast as a well numbered one?
80^3 Kuhn box (3.07 M tets, 393 MB of records > 2x L2 -> binned variant), 8 M particles."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pumiumtally_b200.mesh import kuhn_box
from pumiumtally_b200.tally import PumiTally
from pumiumtally_b200.workload import SyntheticWorkload

cells, n = (80, 80, 80), 8_000_000
coords, t2v = kuhn_box(*cells)
rng = np.random.default_rng(1)
for mode in ("as_generated", "shuffled_numbering"):
    t = t2v if mode == "as_generated" else t2v[rng.permutation(len(t2v))]
    eng = PumiTally.from_arrays(coords, t, n, device=0)
    wl = SyntheticWorkload(box=tuple(float(c) for c in cells), num_particles=n, backend="torch", device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    init = wl.initial_positions().contiguous()
    eng.copy_initial_position_device(init.data_ptr(), s)
    times = []
    for k in range(4):
        o, d, f, w = (x.contiguous() for x in wl.next_step())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.move_device(o.data_ptr(), d.data_ptr(), f.data_ptr(), w.data_ptr(), s); e1.record()
        torch.cuda.synchronize(); times.append(e0.elapsed_time(e1))
    st = eng.stats()
    print(json.dumps({"mode": mode, "variant": eng.get_option("variant"), "ms_per_move": round(float(np.median(times[1:])), 3),
                      "segments_per_move": st["segments"] // 4, "flux_sum": float(eng.flux.sum())}), flush=True)
    del eng
