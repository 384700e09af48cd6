"""Experiment: how fast is the walk when particle index order is spatially sorted?
(upper bound for what a per-step binning pass can buy)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pumiumtally_b200.tally import PumiTally
from pumiumtally_b200.workload import CONFIGS, SyntheticWorkload

cfg = CONFIGS["c2"]; cells = cfg["cells"]; n = cfg["particles"]
box = tuple(float(c) for c in cells)
dev = torch.device("cuda", 0)
for mode in ("random", "sorted_cell", "sorted_elemish"):
    for variant in (4,):
        wl = SyntheticWorkload(box=box, num_particles=n, mean_length=cfg["mean_length"], backend="torch", device=dev)
        init = wl.initial_positions()
        o, d, f, w = wl.next_step()
        if mode != "random":
            c = init.floor().long()
            if mode == "sorted_cell":
                key = (c[:, 2] * cells[1] + c[:, 1]) * cells[0] + c[:, 0]
            else:  # coarser: 4x4x4-cell blocks in z-major order, random inside
                b = c // 4
                key = (b[:, 2] * 14 + b[:, 1]) * 14 + b[:, 0]
            perm = torch.argsort(key)
            init, o, d, f, w = init[perm].contiguous(), o[perm].contiguous(), d[perm].contiguous(), f[perm].contiguous(), w[perm].contiguous()
        eng = PumiTally.from_spec(f"box:{cells[0]},{cells[1]},{cells[2]}", n, device=0)
        eng.set_option("variant", variant)
        s = torch.cuda.current_stream().cuda_stream
        eng.copy_initial_position_device(init.contiguous().data_ptr(), s)
        torch.cuda.synchronize()
        st0 = eng.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.move_device(o.data_ptr(), d.data_ptr(), f.data_ptr(), w.data_ptr(), s)
        e1.record(); torch.cuda.synchronize()
        st1 = eng.stats()
        segs = st1["segments"] - st0["segments"]
        ms = e0.elapsed_time(e1)
        print(json.dumps({"mode": mode, "variant": variant, "ms": ms, "segments": segs, "gseg_s": segs / ms / 1e6,
                          "frac": (128 * segs + 89 * (st1["tracks"] - st0["tracks"])) / (ms * 1e-3) / 6574.1e9}), flush=True)
        del eng
