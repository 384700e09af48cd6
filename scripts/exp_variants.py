"""Kernel-variant sweep on device-resident batches: ms per move and segments/s per variant.
Usage: [PUMITALLY_LIB=pumiumtally_b200/lib/libpumitally_exp.so] python scripts/exp_variants.py <config> <particles|0> v[:block[:opt=val,opt=val]] ..."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pumiumtally_b200.tally import PumiTally
from pumiumtally_b200.workload import CONFIGS, SyntheticWorkload

cfg_name = sys.argv[1]
cfg = CONFIGS[cfg_name]; cells = cfg["cells"]
n = int(sys.argv[2]) or cfg["particles"]
box = tuple(float(c) for c in cells)
dev = torch.device("cuda", 0)
steps, warm = int(os.environ.get("EXP_STEPS", 8)), 3
for spec in sys.argv[3:]:
    parts = spec.split(":")
    v, b, opts = parts[0], (parts[1] if len(parts) > 1 else ""), (parts[2] if len(parts) > 2 else "")
    wl = SyntheticWorkload(box=box, num_particles=n, mean_length=cfg["mean_length"], mu_min=cfg["mu_min"], backend="torch", device=dev)
    init = wl.initial_positions().contiguous()
    eng = PumiTally.from_spec(f"box:{cells[0]},{cells[1]},{cells[2]}", n, device=0)
    try:
        eng.set_option("variant", int(v))
    except ValueError:
        print(json.dumps({"variant": spec, "error": "not in this library"})); continue
    eng.set_option("autotune", 0)
    if b: eng.set_option("block", int(b))
    for kv in filter(None, opts.split(",")):
        k_, v_ = kv.split("="); eng.set_option(k_, int(v_))
    stream = torch.cuda.current_stream().cuda_stream
    eng.copy_initial_position_device(init.data_ptr(), stream)
    ms = []
    for k in range(warm + steps):
        o, d, f, w = (x.contiguous() for x in wl.next_step())
        torch.cuda.synchronize()
        s0 = eng.stats()
        eng.move_device(o.data_ptr(), d.data_ptr(), f.data_ptr(), w.data_ptr(), stream)
        s1 = eng.stats()
        if k >= warm: ms.append((s1["kernel_ms"] - s0["kernel_ms"], s1["segments"] - s0["segments"]))
    t = np.median([m for m, _ in ms]); segs = np.median([s for _, s in ms])
    print(json.dumps({"config": cfg_name, "particles": n, "variant": spec, "ms_per_move": round(float(t), 3), "gseg_s": round(segs / t / 1e6, 2),
                      "flux_sum": float(eng.flux.sum()), "lost": eng.stats()["lost"]}), flush=True)
    del eng
