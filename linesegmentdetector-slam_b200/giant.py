"""One map tiled over several GPUs (BASELINE configs[4]): row bands of the stencil stage, NCCL exchange of the
bands over NVLink, max-all-reduce of the largest gradient, the ordering + region stages on the assembled planes.

Partition.  The scaled image is cut into bands of whole tile rows (32 scaled rows each), `shard_range` over the tile rows.
Rank r holds the whole source map on its GPU — the Gaussian needs source rows floor(y/0.3+0.5) +- 8 around its band
(LSD/myLSD.cpp:460,469), all of which it reads from its own copy, so no halo travels — and runs the stencil stage on its band.
Exchange.  (1) `all_reduce(MAX)` of maxGrad: the bins are quantised against the GLOBAL maximum (LSD/myLSD.cpp:179); (2) every
band of the six row-major planes the stencil writes (mag, deg, (cos,sin), state, ban bits, non-zero bits) is broadcast from its
owner into the same rows of the other ranks' planes — an in-place all-gather, 36.25 bytes per scaled pixel.
Regions.  The seed loop is ONE sequential chain over the map (LSD/myLSD.cpp:218-272): it runs on rank 0 over the assembled planes,
so a region that crosses band borders needs no merge, and logNT / regThre are the global ones (:207-208) by construction.  The
stencil and ordering stages are what the tiling spreads; the region stage's time does not shrink with the GPU count (it is
bound by the latency of one dependent chain)."""
import ctypes as C

import numpy as np

from . import Batch, lib
from .shard import shard_range


class _DevView:
    """a raw device range as a __cuda_array_interface__ object (so that torch can wrap it without copying)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def band_planes(batch):
    """[(device pointer, bytes per scaled row)] of the planes the stencil stage writes, tile rows, scaled rows per tile row"""
    ptrs = (C.c_void_p * 8)(); rb = (C.c_longlong * 8)(); n = C.c_int(0); tr = C.c_int(0); rpt = C.c_int(0)
    batch.ctx.check(lib().lsdb_batch_band_planes(batch.h, 8, ptrs, rb, C.byref(n), C.byref(tr), C.byref(rpt)), "lsdb_batch_band_planes")
    return [(int(ptrs[i]), int(rb[i])) for i in range(n.value)], tr.value, rpt.value


def run_stencil_rows(batch, t0, t1):
    batch.ctx.check(lib().lsdb_batch_run_stencil_rows(batch.h, int(t0), int(t1)), "lsdb_batch_run_stencil_rows")


def max_grad(batch, value=None):
    v = C.c_double(0.0 if value is None else float(value))
    batch.ctx.check(lib().lsdb_batch_max_grad(batch.h, 0 if value is None else 1, C.byref(v)), "lsdb_batch_max_grad")
    return v.value


def run_regions(batch):
    batch.ctx.check(lib().lsdb_batch_run_regions(batch.h), "lsdb_batch_run_regions")


def band_rows(tile_rows, rows_per_tile, H, rank, world):
    """(first tile row, tile rows, first scaled row, one-past-last scaled row) of rank's band"""
    t0, cnt = shard_range(tile_rows, rank, world)
    return t0, cnt, t0 * rows_per_tile, min((t0 + cnt) * rows_per_tile, H)


def lsd_tiled(ctx, map_u8, rank, world, max_lines=65536, want_rects=True):
    """myLineSegmentDetector of ONE map over `world` GPUs (one process per GPU, torch.distributed initialised with NCCL).
    Returns (result dict on rank 0 / None elsewhere, info dict with stage times and the exchange volume)."""
    import torch
    import torch.distributed as dist
    rows, cols = map_u8.shape
    b = Batch(ctx, [(cols, rows)], max_lines=max_lines)
    b.upload([map_u8])
    planes, tile_rows, rpt = band_planes(b)
    H = b.scaled(0)[1]
    stream = torch.cuda.current_stream()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    ev[0].record(stream)
    t0, cnt, y0, y1 = band_rows(tile_rows, rpt, H, rank, world)
    run_stencil_rows(b, t0, t0 + cnt)
    # the library launches on the context's stream, NCCL on torch's: unless the caller made them the same stream (pass
    # torch.cuda.current_stream().cuda_stream of a NON-default stream to Context), order them through the host
    torch.cuda.synchronize()
    ev[1].record(stream)
    sent = 0
    if world > 1:
        g = torch.tensor([max_grad(b)], dtype=torch.float64, device="cuda")
        dist.all_reduce(g, op=dist.ReduceOp.MAX)                     # the global maxGrad, before any binning (LSD/myLSD.cpp:179)
        max_grad(b, float(g.item()))
        for ptr, rb in planes:
            view = torch.as_tensor(_DevView(ptr, rb * H), device="cuda")
            for r in range(world):
                _, _, ry0, ry1 = band_rows(tile_rows, rpt, H, r, world)
                if ry1 > ry0:
                    dist.broadcast(view[ry0 * rb:ry1 * rb], src=r)   # in-place all-gather of the row bands over NVLink
                    if r == rank:
                        sent += (ry1 - ry0) * rb
    torch.cuda.synchronize()
    ev[2].record(stream)
    out = None
    if rank == 0:
        run_regions(b)
        out = b.download(want_rects=want_rects)
    ev[3].record(stream)
    torch.cuda.synchronize()
    info = dict(stencil_ms=ev[0].elapsed_time(ev[1]), exchange_ms=ev[1].elapsed_time(ev[2]), regions_ms=ev[2].elapsed_time(ev[3]),
                total_ms=ev[0].elapsed_time(ev[3]), band_rows=(y0, y1), bytes_sent_by_this_rank=int(sent),
                bytes_per_scaled_row=int(sum(rb for _, rb in planes)), scaled_rows=int(H))
    b.close()
    return out, info
