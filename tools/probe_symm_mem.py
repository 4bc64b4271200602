"""2-GPU probe: does torch symmetric memory (CUDA VMM peer mapping + device-side signal barriers) work on this stack, can
it be captured in a CUDA graph, and what does one halo-sized exchange cost next to NCCL send/recv?
    torchrun --nproc-per-node 2 tools/probe_symm_mem.py"""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as sm


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    n = 1 << 20  # 4 MB of floats: one full-resolution halo row set
    box = sm.empty(n, dtype=torch.float32, device=dev)
    hdl = sm.rendezvous(box, dist.group.WORLD)
    peer = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
    src, dst = torch.full((n,), float(rank + 1), device=dev), torch.zeros(n, device=dev)

    def exchange():
        box.copy_(src)
        hdl.barrier(channel=0)
        dst.copy_(peer)
        hdl.barrier(channel=1)

    exchange()
    torch.cuda.synchronize()
    want = float((rank + 1) % world + 1)
    assert float(dst[0]) == want and float(dst[-1]) == want, (float(dst[0]), want)
    print("rank %d: eager peer exchange ok" % rank, flush=True)

    def timeit(fn, iters=200):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3, (time.perf_counter() - t0) / iters * 1e6

    t_symm = timeit(exchange)

    buf = torch.empty(n, device=dev)

    def nccl_exchange():
        ops = [dist.P2POp(dist.isend, src, (rank + 1) % world), dist.P2POp(dist.irecv, buf, (rank - 1) % world)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()

    t_nccl = timeit(nccl_exchange)
    sums = torch.ones(256, device=dev, dtype=torch.float64)
    t_ar = timeit(lambda: dist.all_reduce(sums))
    if rank == 0:
        print("4 MB exchange: symm-mem copy+2 barriers %.1f us GPU / %.1f us wall ; NCCL send/recv %.1f / %.1f us ; "
              "all_reduce(256 doubles) %.1f / %.1f us" % (t_symm + t_nccl + t_ar), flush=True)

    # CUDA graph capture of the symmetric-memory exchange
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        exchange()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    dist.barrier()
    with torch.cuda.graph(g):
        exchange()
    for k in range(3):
        src.fill_(float(10 * k + rank + 1))
        g.replay()
        torch.cuda.synchronize()
        want = float(10 * k + (rank + 1) % world + 1)
        assert float(dst[0]) == want, (k, float(dst[0]), want)
    t_graph = timeit(g.replay)
    if rank == 0:
        print("graph replay of the exchange ok: %.1f us GPU / %.1f us wall" % t_graph, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
