"""Run under torchrun (>= 2 GPUs): the multi-view step with the peer-memory exchange (peer.PeerExchange, CUDA IPC + this
library's kernels) against the same step with the NCCL exchange (parallel.MultiViewExchange).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_peer_exchange.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from artdeco_b200 import synthetic
    from artdeco_b200.multiview import MultiViewStep
    from artdeco_b200.parallel import views_for_rank
    from artdeco_b200.peer import PeerExchange

    N, W, H, n_views = 40000, 640, 368, 8
    sc = synthetic.raster_scene(N, seed=3)
    cams = [synthetic.camera(W, H, view=0.5 + v) for v in range(n_views)]
    mine = views_for_rank(n_views, world, rank)
    Vs = torch.stack([cams[v][0] for v in mine]).to(dev)
    Ks = torch.stack([cams[v][1] for v in mine]).to(dev)
    params = {k: sc[k].to(dev) for k in ("means", "quats", "scales", "opacities", "sh")}
    g = torch.Generator().manual_seed(11)
    vc_all, va_all = torch.randn(n_views, H, W, 4, generator=g), torch.randn(n_views, H, W, generator=g)
    results = {}
    for kind in ("nccl", "peer"):
        for graph in (False, True):
            eng = MultiViewStep(params, Vs, Ks, W, H, world=world, graph=graph, exchange_kind=kind)
            if kind == "peer":
                assert isinstance(eng.exchange, PeerExchange), f"peer mapping unavailable: got {type(eng.exchange).__name__}"
            eng.set_upstream(vc_all[mine].to(dev), va_all[mine].to(dev))
            for _ in range(3):                       # several steps: both parities of the colour table, counters advancing
                grads = eng.step()
            torch.cuda.synchronize()
            eng.check_overflow()
            if kind == "peer":
                eng.exchange.check()
            results[(kind, graph)] = {k: grads[k].clone() for k in ("v_means", "v_quats", "v_scales", "v_opac", "v_sh")}
            dist.barrier()
    ok = True
    ref = results[("nccl", False)]
    for key, res in results.items():
        for k, v in res.items():
            scale = ref[k].abs().max().clamp_min(1e-12)
            err = float((v - ref[k]).abs().max() / scale)
            if not err < 2e-5:
                ok = False
                print(f"rank {rank}: {key} {k}: rel err {err:.3e}", flush=True)
    # every rank must hold the same reduced gradients (peer path: bit-identical by construction)
    mine_flat = torch.cat([results[("peer", True)][k].reshape(-1) for k in ("v_means", "v_quats", "v_scales", "v_opac")])
    other = [torch.empty_like(mine_flat) for _ in range(world)]
    dist.all_gather(other, mine_flat)
    same = all(torch.equal(o, other[0]) for o in other)
    flag = torch.tensor([int(ok and same)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("PEER-EXCHANGE-OK" if int(flag) else f"PEER-EXCHANGE-MISMATCH (ok={ok}, identical across ranks={same})", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
