"""Under torchrun: the config-4 partitioned fusion of bench.py for several slab widths / axes (load balance against halo size)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for slab in [int(x) for x in (sys.argv[1:] or ["8", "4", "2", "1"])]:
    bench.SLAB4 = slab
    c4 = bench.bench_config4(local, rank, world, 20, 5, False)
    if rank == 0:
        p = c4["partitioned"]
        print(json.dumps({"slab": slab, "ms_per_frame": p["ms_per_frame"], "fps": 1e3 / p["ms_per_frame"], "e2e_fps": 1e3 / p["e2e_ms_per_frame"],
                          "integrate_ms_max": c4.get("integrate_ms_max_over_ranks"), "select_ms_max": c4.get("select_ms_max_over_ranks"),
                          "rank0_frame_cubes": p["frame_cubes"], "boundary_cubes": c4.get("boundary_cubes_exchanged"),
                          "halo_ms": c4.get("halo_exchange_ms_max_over_ranks"), "mesh_vertices": c4.get("mesh_vertices_total")}), flush=True)
dist.destroy_process_group()
