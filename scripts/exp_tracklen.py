"""Streaming (8) vs packed/sorted (24) walk as a function of track length (segments per track), c2 mesh."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pumiumtally_b200.tally import PumiTally
from pumiumtally_b200.workload import CONFIGS, SyntheticWorkload

cells = CONFIGS["c2"]["cells"]; box = tuple(float(c) for c in cells)
dev = torch.device("cuda", 0)
for mean_length, n in ((3.0, 10_000_000), (8.0, 5_000_000), (15.0, 3_000_000), (30.0, 2_000_000), (60.0, 1_000_000)):
    out = {"mean_length": mean_length, "particles": n}
    for variant in (8, 24):
        wl = SyntheticWorkload(box=box, num_particles=n, mean_length=mean_length, backend="torch", device=dev)
        eng = PumiTally.from_spec(f"box:{cells[0]},{cells[1]},{cells[2]}", n, device=0)
        eng.set_option("variant", variant)
        s = torch.cuda.current_stream().cuda_stream
        eng.copy_initial_position_device(wl.initial_positions().contiguous().data_ptr(), s)
        ms = []
        for step in range(5):
            o, d, f, w = (x.contiguous() for x in wl.next_step())
            st0 = eng.stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.move_device(o.data_ptr(), d.data_ptr(), f.data_ptr(), w.data_ptr(), s); e1.record()
            torch.cuda.synchronize()
            st1 = eng.stats()
            ms.append(e0.elapsed_time(e1))
        segs, tracks = st1["segments"] - st0["segments"], st1["tracks"] - st0["tracks"]
        out[f"v{variant}_ms"] = round(sorted(ms[1:])[len(ms[1:]) // 2], 3)
        out["seg_per_track"] = round(segs / tracks, 1)
        del eng
    out["packed_over_streaming"] = round(out["v24_ms"] / out["v8_ms"], 3)
    print(json.dumps(out), flush=True)
