"""Two-rank probe of the peer-memory halo kernels (run under torchrun on a 2-GPU box):
symmetric-memory allocation + rendezvous, fegnn_halo_push / fegnn_halo_reduce_push into the peer, barrier,
verification, and the time of (push + barrier) against an NCCL all-to-all of the same rows."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm
    from fastegnn_b200 import _lib as L
    lib = L.lib
    st = torch.cuda.current_stream().cuda_stream
    N, Hn = 20000, 6000                      # owned rows, halo rows
    Nl = N + Hn
    Q = symm.empty(Nl, 64, dtype=torch.float32, device=dev)
    x = symm.empty(Nl, 3, dtype=torch.float32, device=dev)
    hq = symm.rendezvous(Q, dist.group.WORLD)
    hx = symm.rendezvous(x, dist.group.WORLD)
    print(f"[{rank}] buffer_ptrs {[hex(p) for p in hq.buffer_ptrs]} multicast {hq.has_multicast_support}", flush=True)
    peer = 1 - rank
    g = torch.Generator(device="cpu").manual_seed(rank)
    Q.zero_(); x.zero_()
    Q[:N] = torch.randn(N, 64, generator=g).to(dev)
    x[:N] = torch.randn(N, 3, generator=g).to(dev)
    # send my rows src[k] into the peer's halo row N + k
    src = torch.randperm(N, generator=torch.Generator().manual_seed(7 + rank))[:Hn].to(torch.int32).to(dev)
    k = torch.arange(Hn, dtype=torch.int64, device=dev)
    dst_q = (hq.buffer_ptrs[peer] + (N + k) * 256)
    dst_x = (hx.buffer_ptrs[peer] + (N + k) * 12)
    hq.barrier(channel=0)
    L.check(lib.fegnn_halo_push(Hn, L.ptr(src), L.ptr(dst_q), L.ptr(dst_x), L.ptr(Q), L.ptr(x), st), "halo_push")
    hq.barrier(channel=0)
    torch.cuda.synchronize()
    # what the peer sent me: its rows src_peer
    gp = torch.Generator(device="cpu").manual_seed(peer)
    Qp = torch.randn(N, 64, generator=gp).to(dev)
    xp = torch.randn(N, 3, generator=gp).to(dev)
    srcp = torch.randperm(N, generator=torch.Generator().manual_seed(7 + peer))[:Hn].to(dev).long()
    ok1 = bool(torch.equal(Q[N:], Qp[srcp]) and torch.equal(x[N:], xp[srcp]))
    print(f"[{rank}] push correct: {ok1}", flush=True)
    # reverse: add my halo rows (gradients) into the owner's rows
    gQ = symm.empty(Nl, 64, dtype=torch.float32, device=dev)
    gx = symm.empty(Nl, 3, dtype=torch.float32, device=dev)
    hgq = symm.rendezvous(gQ, dist.group.WORLD)
    hgx = symm.rendezvous(gx, dist.group.WORLD)
    gQ.fill_(1.0 + rank); gx.fill_(0.5 + rank)
    dq = (hgq.buffer_ptrs[peer] + srcp * 256)      # my halo row k is the peer's row srcp[k]
    dx = (hgx.buffer_ptrs[peer] + srcp * 12)
    hgq.barrier(channel=0)
    L.check(lib.fegnn_halo_reduce_push(Hn, N, L.ptr(dq), L.ptr(dx), L.ptr(gQ), L.ptr(gx), st), "halo_reduce_push")
    hgq.barrier(channel=0)
    torch.cuda.synchronize()
    mine = src.long()                               # my rows that the peer holds as halo: each got + (1 + peer)
    exp = torch.full((N,), 1.0 + rank, device=dev)
    exp[mine] += 1.0 + peer
    ok2 = bool(torch.equal(gQ[:N, 0], exp) and torch.equal(gQ[:N, 63], exp))
    print(f"[{rank}] reduce_push correct: {ok2}", flush=True)
    # timing: (push + barrier) vs index_select + all_to_all_single + copy
    def t_p2p():
        L.check(lib.fegnn_halo_push(Hn, L.ptr(src), L.ptr(dst_q), L.ptr(dst_x), L.ptr(Q), L.ptr(x), st), "halo_push")
        hq.barrier(channel=0)
    def t_nccl():
        send = torch.cat([Q.index_select(0, src.long()), x.index_select(0, src.long())], dim=1)
        recv = torch.empty(Hn, 67, device=dev)
        cnt = [0, 0]; cnt[peer] = Hn
        dist.all_to_all_single(recv, send, cnt, cnt)
        Q[N:] = recv[:, :64]; x[N:] = recv[:, 64:]
    for name, fn in (("p2p push+barrier", t_p2p), ("nccl pack+all_to_all+unpack", t_nccl)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(50):
            fn()
        e.record(); torch.cuda.synchronize()
        print(f"[{rank}] {name}: {s.elapsed_time(e) / 50 * 1e3:.1f} us per exchange of {Hn} rows", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if (ok1 and ok2) else 1)


if __name__ == "__main__":
    main()
