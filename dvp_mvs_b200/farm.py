"""Per-view sharding of a scene across the GPUs of one box.

One reference view = one GPU job.  Within a pass every view is independent given the previous pass's maps, so
views are dealt to ranks and processed with NO data-path collective; the only communication is the exchange of
per-view results at the end of a pass (depth maps feed the next pass's geometric consistency), done with
torch.distributed (NCCL on GPUs, gloo in the CPU tests).  The reference processes views sequentially in one
process and exchanges maps through files (main.cpp:486-507, APD.cpp:1150-1158); a parallel pass is therefore
Jacobi-ordered (all views see pass k-1 maps), which is stated wherever end-to-end results are compared.

One process per GPU, launched with torchrun; rank r owns device LOCAL_RANK.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Sequence

import numpy as np


def partition(num_views: int, world: int, rank: int, costs: Sequence[float] | None = None) -> List[int]:
    """Views owned by `rank`.  Equal-cost views: round-robin.  With per-view cost estimates (e.g. pixel count
    times number of source views, or the WEAK-pixel count of the previous pass): longest-processing-time-first
    greedy, deterministic (ties broken by view id) so that every rank computes the same assignment."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    if costs is None:
        return list(range(rank, num_views, world))
    if len(costs) != num_views:
        raise ValueError("costs must have one entry per view")
    order = sorted(range(num_views), key=lambda v: (-float(costs[v]), v))
    load = [0.0] * world
    owner: Dict[int, int] = {}
    for v in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[v] = r
        load[r] += float(costs[v])
    return sorted(v for v, r in owner.items() if r == rank)


def run_pass(num_views: int, process_view: Callable[[int], Dict[str, np.ndarray]], costs: Sequence[float] | None = None,
             gather: bool = True) -> Dict[int, Dict[str, np.ndarray]]:
    """Process this rank's share of the views with `process_view(view_id) -> {name: array}` and (optionally)
    all-gather the results so every rank holds every view's maps for the next pass.
    Works without torch.distributed being initialised (single process)."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = partition(num_views, world, rank, costs)
    local = {v: process_view(v) for v in mine}
    if world == 1 or not gather:
        return local
    import torch
    backend = dist.get_backend()
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if backend == "nccl" else torch.device("cpu")
    # every view has the same set of maps; the owner broadcasts each of them (no reduction, no staging through files)
    out: Dict[int, Dict[str, np.ndarray]] = {}
    for v in range(num_views):
        owner = next(r for r in range(world) if v in partition(num_views, world, r, costs))
        meta = [None]
        if rank == owner:
            meta = [{k: (a.shape, str(a.dtype)) for k, a in local[v].items()}]
        dist.broadcast_object_list(meta, src=owner)
        maps = {}
        for k, (shape, dtype) in meta[0].items():
            if rank == owner:
                t = torch.from_numpy(np.ascontiguousarray(local[v][k]).view(np.uint8).reshape(-1)).to(device)
            else:
                t = torch.empty(int(np.prod(shape)) * np.dtype(dtype).itemsize, dtype=torch.uint8, device=device)
            dist.broadcast(t, src=owner)
            maps[k] = t.cpu().numpy().view(dtype).reshape(shape)
        out[v] = maps
    return out


def run_scene_schedule(scene, num_views: int, num_levels: int, seed: int, costs: Sequence[float] | None = None) -> Dict[int, int]:
    """The multi-scale schedule (main.cpp:449-511) with the views of every pass dealt to the ranks.

    `scene` holds every view's pyramid on this rank's GPU (dvp_mvs_b200.Scene, or anything with `run_view(view, level,
    pass, seed)` and `depth_tensor(view, level, owned) -> torch tensor`).  Each rank runs its own views of a pass in
    index order; the one exchange step per pass is ONE all-gather of the ranks' fresh depth maps between the scenes' device
    buffers (`exchange_depths`: NCCL; gloo in the CPU tests) — depth is all a view needs from its sources.  Within a
    pass a rank sees its own views' fresh depths and the other ranks' previous-pass depths (block Gauss-Seidel); with
    one rank this is exactly the reference's sequential order.  Returns {view: owner rank}."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    owner = {v: r for r in range(world) for v in partition(num_views, world, r, costs)}
    mine = [v for v in range(num_views) if owner[v] == rank]
    it = 0
    for level in range(num_levels):
        for pass_ in range(4):
            for v in mine:
                scene.run_view(v, level, pass_, seed + 1000 * it + v)   # same seeds as dvp_scene_run
            if world > 1:
                exchange_depths(scene, level, owner, rank, world)
            it += 1
    return owner


def exchange_depths(scene, level: int, owner: Dict[int, int], rank: int, world: int):
    """The one exchange step of a farmed pass: ONE all-gather of every rank's freshly computed depth maps
    (all views of a level have the same size; ranks owning fewer views pad with an unused slot), then the remote maps are
    copied into the scene's buffers.  Synchronisation: `run_view` returns only after its stream has drained, so the maps
    packed here are complete; the collective and the unpacking copies run on torch's streams, which the scene's own
    non-blocking stream does not order against — the device is therefore synchronised before the next pass may read a
    remote depth map or overwrite a local one that NCCL is still sending."""
    import torch
    import torch.distributed as dist
    views_of = {r: sorted(v for v, o in owner.items() if o == r) for r in range(world)}
    slots = max(1, max(len(v) for v in views_of.values()))
    any_view = next(iter(sorted(owner)))
    shape = tuple(scene.depth_tensor(any_view, level, owner[any_view] == rank).shape)
    first = scene.depth_tensor(views_of[rank][0], level, True) if views_of[rank] else scene.depth_tensor(any_view, level, False)
    send = torch.zeros((slots,) + shape, dtype=torch.float32, device=first.device)
    for i, v in enumerate(views_of[rank]):
        send[i].copy_(scene.depth_tensor(v, level, True))
    flat = torch.empty((world * slots,) + shape, dtype=torch.float32, device=first.device)   # rank-major concatenation
    dist.all_gather_into_tensor(flat, send)
    recv = flat.view((world, slots) + shape)
    for r in range(world):
        if r == rank:
            continue
        for i, v in enumerate(views_of[r]):
            scene.depth_tensor(v, level, False).copy_(recv[r, i])
    if first.is_cuda:
        torch.cuda.synchronize(first.device)


def fuse_farmed_scene(scene, owner: Dict[int, int], views_static: Sequence[dict], fuse_rank: int = 0, mode: int = 0, make_fusion=None):
    """Row N3 after a farmed schedule: the finished plane and pixel-state maps of every view meet on `fuse_rank`, which
    fuses them (RunFusion, APD.cpp:1809-1960; `mode` 1 / 2 for the T&T variants).  Fusion costs well under a millisecond
    of device time per megapixel view, so it runs on ONE GPU ("replicas only"); the exchange is a broadcast
    of each view's maps from its owner — the in-memory form of the depths.dmb / APD_normals.dmb / weak.bin files the
    reference's RunFusion reads (APD.cpp:1851-1858).

    `scene.get_view(v) -> (planes [h,w,4], weak [h,w], ...)` on the owner; `views_static[v]` = dict(camera=<camera at the
    map size>, image=<[h,w,3] uint8>, src_views=[...], block=None) known to every rank.  Returns the point list
    [n, 6] on `fuse_rank`, None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    num_views = len(views_static)
    device = torch.device("cpu")
    if world > 1 and dist.get_backend() == "nccl":
        device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    gathered = []
    for v in range(num_views):
        src = owner[v] if world > 1 else 0
        planes = weak = None
        if rank == src:
            got = scene.get_view(v)
            planes = np.ascontiguousarray(got[0], np.float32); weak = np.ascontiguousarray(got[1], np.uint8)
        if world > 1:
            shape = [tuple(planes.shape[:2]) if rank == src else None]
            dist.broadcast_object_list(shape, src=src)
            h, w = shape[0]
            tp = torch.from_numpy(planes).to(device) if rank == src else torch.empty((h, w, 4), dtype=torch.float32, device=device)
            tw = torch.from_numpy(weak).to(device) if rank == src else torch.empty((h, w), dtype=torch.uint8, device=device)
            dist.broadcast(tp, src=src); dist.broadcast(tw, src=src)
            if rank == fuse_rank and rank != src:
                planes, weak = tp.cpu().numpy(), tw.cpu().numpy()
        if rank == fuse_rank:
            st = views_static[v]
            gathered.append(dict(camera=st["camera"], planes=planes, image=st["image"], weak=weak, src_views=list(st["src_views"]),
                                 block=st.get("block")))
    if rank != fuse_rank:
        return None
    if make_fusion is None:
        from . import Fusion
        make_fusion = lambda views: Fusion(views, device=int(os.environ.get("LOCAL_RANK", "0")))
    f = make_fusion(gathered)
    if mode:
        f.set_mode(mode)
    points, _ = f.run()
    return points


def make_gpu_view_processor(scenes, params, iters_hint: int | None = None):
    """process_view for synthetic scenes: one Engine per rank, reused across its views."""
    from . import Engine
    cache = {}

    def process(v: int):
        sc = scenes[v]
        key = (sc.width, sc.height, sc.num_src)
        if key not in cache:
            cache[key] = Engine(sc.width, sc.height, sc.num_src, params, device=int(os.environ.get("LOCAL_RANK", "0")))
        e = cache[key]
        e.upload(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, seed=1000 + v, params=params)
        e.run()
        planes, weak, sel, rad = e.download()
        return dict(depth=np.ascontiguousarray(planes[..., 3]), normal=np.ascontiguousarray(planes[..., :3]), weak=weak, selected=sel, radius=rad)
    return process
