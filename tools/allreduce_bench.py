"""SUM all-reduce of the data-parallel step's gradient buffer: this library's peer-memory kernel (zs_allreduce_sum_peer)
against torch.distributed.all_reduce (NCCL), per call, on N ranks of one box.  Run under torchrun:
    python -m torch.distributed.run --nproc-per-node N tools/allreduce_bench.py
Rank 0 prints one JSON line per size."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zhusuan-pytorch_b200"))
import torch
import torch.distributed as dist
import zhusuan.distributed as zd

os.environ.setdefault("NCCL_DEBUG", "WARN")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def timed(fn, reps=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


for n in (4, 663784, 1346868):
    peer = zd._PeerBuffer(n, dev, torch.float32, None)
    peer.flat.fill_(float(rank + 1))
    torch.cuda.synchronize(); dist.barrier()
    peer.all_reduce(0, (n + 3) // 4 * 4)
    torch.cuda.synchronize()
    ok = bool((peer.flat == world * (world + 1) / 2).all())
    res = {"floats": n, "world": world, "peer_correct": ok, "variant": peer.variant}
    if peer.mc_ptr:  # also the peer loads / stores form on the same buffer
        mc, peer.mc_ptr = peer.mc_ptr, 0
        peer.flat.fill_(float(rank + 1))
        torch.cuda.synchronize(); dist.barrier()
        peer.all_reduce(0, (n + 3) // 4 * 4)
        torch.cuda.synchronize()
        res["p2p_correct"] = bool((peer.flat == world * (world + 1) / 2).all())
        res["p2p_us"] = round(timed(lambda: peer.all_reduce(0, (n + 3) // 4 * 4)), 2)
        peer.mc_ptr = mc
        res["nvls_us"] = round(timed(lambda: peer.all_reduce(0, (n + 3) // 4 * 4)), 2)
        for ctas in (8, 16, 32, 64):
            import zhusuan._backend as be
            res["nvls_us_ctas%d" % ctas] = round(timed(lambda: be.allreduce_sum_nvls(peer.mc_ptr, peer.buf_ptrs[peer.rank], peer.flag_ptrs, peer.rank, 0, (n + 3) // 4 * 4, 0, dev, ctas)), 2)
    for ctas in (16, 32, 64):
        import zhusuan._backend as be
        res["peer_us_ctas%d" % ctas] = round(timed(lambda: be.allreduce_sum_peer(peer.buf_ptrs, peer.flag_ptrs, peer.rank, 0,
                                                                                   (n + 3) // 4 * 4, 0, dev, ctas)), 2)
    # the kernel itself, without the host's launch cost: 10 launches per graph replay
    gp = torch.cuda.CUDAGraph()
    sp = torch.cuda.Stream()
    sp.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(sp):
        with torch.cuda.graph(gp, stream=sp):
            for _ in range(10):
                peer.all_reduce(0, (n + 3) // 4 * 4, flag_set=1)
    torch.cuda.current_stream().wait_stream(sp)
    res["peer_us_in_graph"] = round(timed(gp.replay, reps=30, warm=3) / 10, 2)
    t = torch.ones(n, device=dev)
    res["nccl_us"] = round(timed(lambda: dist.all_reduce(t)), 2)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        dist.all_reduce(t)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(10):
                dist.all_reduce(t)
    torch.cuda.current_stream().wait_stream(s)
    res["nccl_us_in_graph"] = round(timed(g.replay, reps=30, warm=3) / 10, 2)
    if rank == 0:
        print(json.dumps(res), flush=True)
    keep = globals().setdefault("_keep", [])
    keep.append((peer, gp, g))  # peer-mapped buffers stay alive until every rank is done
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
os._exit(0)  # skip the teardown of symmetric memory / NCCL (the previous version of this tool hung there)
