#!/usr/bin/env python3
"""One synthetic map through giant.lsd_tiled on all ranks of a torchrun launch, checked on rank 0 against the same map on one GPU
(and, with --oracle, against the CPU oracle).  torchrun --nproc-per-node N tools/giant_check.py [--size 16384] [--oracle]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--seed", type=int, default=5000)
    ap.add_argument("--oracle", action="store_true")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from __graft_entry__ import load_package
    lsdb = load_package()
    from lsdb200 import giant
    import synth
    m = synth.occupancy_grid(a.size, a.size, seed=a.seed)
    torch.cuda.set_stream(torch.cuda.Stream())                # a real stream: handle 0 would make the library create its own
    ctx = lsdb.Context(local, torch.cuda.current_stream().cuda_stream)
    giant.lsd_tiled(ctx, m, rank, world)                      # warm-up (allocations, NCCL channels)
    out, info = giant.lsd_tiled(ctx, m, rank, world)
    if rank == 0:
        b = lsdb.Batch(ctx, [(a.size, a.size)], max_lines=65536)
        b.upload([m]); b.run(); one = b.download(want_rects=True); st = b.stage_ms(); b.close()
        same = int(out["counts"][0]) == int(one["counts"][0]) and np.array_equal(out["rects"][0], one["rects"][0], equal_nan=True) and \
            np.array_equal(lsdb.lines_to_array(out["lines"][0]), lsdb.lines_to_array(one["lines"][0]), equal_nan=True)
        ok_oracle = None
        if a.oracle:
            import oraclebind
            o = oraclebind.lsd(m, want_maps=False, want_line_im=False, max_lines=65536)
            ok_oracle = int(out["counts"][0]) == o["n"] and np.array_equal(out["rects"][0], o["rects"], equal_nan=True)
        print(json.dumps(dict(size=a.size, n_gpus=world, segments=int(out["counts"][0]), tiled=info, single_gpu_stage_ms=st,
                              equals_single_gpu=bool(same), equals_oracle=ok_oracle)))
        if not same or ok_oracle is False:
            raise SystemExit("TILED != SINGLE")
        print("TILED == SINGLE")
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
