#!/usr/bin/env python
"""Developer bench: ONLY the tp_cfg4 leg of bench.py (cfg4 decode layer step with the fused GEMM + all-reduce) under
torchrun, one JSON line on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29535 \
        tools/bench_tp_step.py [--steps 20 --warmup 5]
"""
import argparse, json, os, sys
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    os.environ["MOJO_BACKEND"] = "b200"
    os.environ.setdefault("MOJO_B200_GAR_TIMEOUT_S", "60")
    import mojo_opset_b200 as m

    def barrier():
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    leg, ok = bench.run_tp_cfg4(m, dev, rank, world, bench.load_peaks(), args, barrier, max_over_ranks)
    if rank == 0:
        keep = {k: leg[k] for k in ("tp", "ms_per_step", "ms_per_step_cublas_nccl", "ms_per_step_no_collective",
                                    "ms_attention_only", "allreduce_exposed_us", "allreduce_exposed_us_cublas_nccl",
                                    "gemm_allreduce_us", "cublas_plus_nccl_us", "decode_gbs_per_rank", "parity")}
        keep["env"] = {k: v for k, v in os.environ.items() if k.startswith("MOJO_B200_")}
        print(json.dumps(keep), flush=True)
    dist.barrier()
    from mojo_opset_b200.backends.b200.operators.compute_with_comm import release_workspaces
    release_workspaces()
    dist.destroy_process_group()
    sys.exit(0 if ok else 3)


if __name__ == "__main__":
    main()
