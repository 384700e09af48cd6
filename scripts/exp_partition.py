"""A/B on N GPUs (torchrun): particle-striped replicas (the engine's default multi-GPU scheme: every
rank holds the whole mesh, one ncclAllReduce of the flux per batch) against the spatial partition
(pumiumtally_b200/partition.py: RCB picparts with ghost layers, particles routed to the owner of
their tet, hand-off at picpart boundaries, ghost-only tally exchange).  Same mesh, same particles.
Usage: torchrun --nproc-per-node N scripts/exp_partition.py [config] [particles_total] [layers] [steps]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from pumiumtally_b200.distributed import broadcast_unique_id
from pumiumtally_b200.mesh import kuhn_box
from pumiumtally_b200.partition import PartitionedTally
from pumiumtally_b200.tally import PumiTally
from pumiumtally_b200.workload import CONFIGS, SyntheticWorkload

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "c5"
cfg = CONFIGS[cfg_name]
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
n_total = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["particles"]
layers = int(sys.argv[3]) if len(sys.argv) > 3 else 2
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cells = cfg["cells"]
box = tuple(float(c) for c in cells)
n = n_total // world
coords, t2v = kuhn_box(*cells)
stream = torch.cuda.current_stream().cuda_stream

def workload():
    w = SyntheticWorkload(box=box, num_particles=n, mean_length=cfg["mean_length"], mu_min=cfg["mu_min"], backend="torch", device=dev, id_offset=rank * n)
    return w, w.initial_positions().contiguous()

def sync():
    torch.cuda.synchronize(); dist.barrier()

# ---------------- A: replicas
wl, init = workload()
eng = PumiTally.from_spec(f"box:{cells[0]},{cells[1]},{cells[2]}", n, device=local)
eng.comm_init(rank, world, broadcast_unique_id(dist, PumiTally.nccl_unique_id, device=dev))
eng.copy_initial_position_device(init.data_ptr(), stream)
times = []
for k in range(steps):
    o, d, f, w = (x.contiguous() for x in wl.next_step())
    sync(); t0 = time.perf_counter()
    eng.move_device(o.data_ptr(), d.data_ptr(), f.data_ptr(), w.data_ptr(), stream)
    sync(); times.append(time.perf_counter() - t0)
sync(); t0 = time.perf_counter()
eng.allreduce_tally()
sync(); t_ar = time.perf_counter() - t0
st = eng.stats()
flux_a = torch.from_numpy(eng.flux).to(dev)
elem_a, pos_a = torch.from_numpy(eng.elem_ids.astype(np.int64)).to(dev), torch.from_numpy(eng.positions).to(dev)
segs = torch.tensor([float(st["segments"])], dtype=torch.float64, device=dev); dist.all_reduce(segs)
tA = torch.tensor([np.median(times[1:])], dtype=torch.float64, device=dev); dist.all_reduce(tA, op=dist.ReduceOp.MAX)
table_a = eng.num_elements * 128
del eng
torch.cuda.empty_cache()

# ---------------- B: spatial partition
wl, init = workload()
t0 = time.perf_counter()
pt = PartitionedTally(coords, t2v, n, dist, dev, layers=layers)
sync(); t_build = time.perf_counter() - t0
pt.CopyInitialPosition(init)
times_b = []
for k in range(steps):
    o, d, f, w = (x.contiguous() for x in wl.next_step())
    pt.profile = k >= 1
    sync(); t0 = time.perf_counter()
    pt.MoveToNextLocation(o, d, f, w)
    sync(); times_b.append(time.perf_counter() - t0)
sync(); t0 = time.perf_counter()
owned = pt.exchange_ghost_tallies()
sync(); t_ghost = time.perf_counter() - t0
flux_b = pt.global_flux()
tB = torch.tensor([np.median(times_b[1:])], dtype=torch.float64, device=dev); dist.all_reduce(tB, op=dist.ReduceOp.MAX)
tol = 1e-6 * flux_a.abs() + 1e-12 * flux_a.sum()
bad = int(((flux_b - flux_a).abs() > tol).sum())
worst = float(((flux_b - flux_a).abs() / (flux_a.abs() + 1e-12 * flux_a.sum())).max())
elem_mismatch = torch.tensor([int((pt.elem_ids != elem_a).sum())], device=dev); dist.all_reduce(elem_mismatch)
pos_err = torch.tensor([float((pt.positions - pos_a).abs().max())], dtype=torch.float64, device=dev); dist.all_reduce(pos_err, op=dist.ReduceOp.MAX)
hand = torch.tensor([pt.stats_handoffs, pt.stats_routed, pt.stats_rounds], dtype=torch.float64, device=dev); dist.all_reduce(hand)
if rank == 0:
    per_move_segs = float(segs) / steps
    tim = {k: round(1e3 * v / max(steps - 1, 1), 3) for k, v in pt.timers.items()}
    print(json.dumps({
        "config": cfg_name, "n_gpus": world, "particles_total": n * world, "tets": len(t2v), "ghost_layers": layers,
        "replicas": {"ms_per_move": round(1e3 * float(tA), 3), "gseg_s": round(per_move_segs / float(tA) / 1e9, 2),
                     "tet_table_MB_per_gpu": round(table_a / 1e6, 1), "allreduce_ms": round(1e3 * t_ar, 3),
                     "allreduce_MB": round(8 * len(t2v) / 1e6, 1)},
        "partition": {"ms_per_move": round(1e3 * float(tB), 3), "gseg_s": round(per_move_segs / float(tB) / 1e9, 2),
                      "tet_table_MB_per_gpu": round(pt.pic.n_local * 128 / 1e6, 1), "owned_tets": pt.pic.n_owned,
                      "local_tets": pt.pic.n_local, "ghost_exchange_ms": round(1e3 * t_ghost, 3),
                      "ghost_exchange_MB": round(8 * (pt.pic.n_local - pt.pic.n_owned) / 1e6, 2),
                      "phases_ms_per_move_rank0": tim, "handoffs_per_move": float(hand[0]) / steps,
                      "routed_per_move": float(hand[1]) / steps, "rounds_per_move": float(hand[2]) / steps / world,
                      "build_s": round(t_build, 1)},
        "parity": {"flux_elements_outside_1e-6": bad, "worst_rel": worst, "parent_element_mismatches": int(elem_mismatch),
                   "max_position_error": float(pos_err)}}), flush=True)
dist.barrier()
dist.destroy_process_group()
