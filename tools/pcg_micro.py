"""Times the Jacobi-PCG iteration alone (fixed iteration count, cold start) on a stand-in lattice.
python tools/pcg_micro.py [workload=standin8x8_brick] [iters=60]      (or under torchrun for the row-sharded form)
Prints one line: PCG_MICRO {json}.  Used for A/B runs of environment switches (KMCB200_PDL, ...)."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
PKG = bench.PKG


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    name = sys.argv[1] if len(sys.argv) > 1 else "standin8x8_brick"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kmc = importlib.import_module(PKG)
    ctx = kmc.Context(local)
    s, _ = bench.build_workload(kmc, name)
    r = bench.pcg_fixed_iterations(kmc, ctx, s, rank, world, dist, iters=iters, solves=4)
    r["env"] = {k: v for k, v in os.environ.items() if k.startswith("KMCB200_")}
    r["frac_of_hbm_peak_per_gpu"] = r["GBs_per_gpu"] / bench.peaks()[0]
    if rank == 0:
        print("PCG_MICRO " + json.dumps(r), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
