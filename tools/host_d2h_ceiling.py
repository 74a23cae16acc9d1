#!/usr/bin/env python
"""How much device->host traffic one box can absorb: N ranks (one per GPU, torchrun) copy from their GPU into pinned host memory AT THE SAME
TIME, back to back, for a few seconds; rank 0 prints per-rank and aggregate GB/s.  This is the ceiling of the decoder's float end-to-end leg
(bench.py `e2e`), which moves 3840 bytes of PCM per mono frame to the host: at N = 8 the ranks share one host memory system.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/host_d2h_ceiling.py [--mb 512] [--seconds 2]
  python tools/host_d2h_ceiling.py            # N = 1
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--write-combined", action="store_true", help="allocate the host target with cudaHostAllocWriteCombined instead of torch's pinned allocator")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = args.mb << 20
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    res = {}
    for kind in (["pinned", "write_combined"] if args.write_combined else ["pinned"]):
        if kind == "pinned":
            h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            hp = h.data_ptr()
        else:
            import ctypes
            rt = ctypes.CDLL("libcudart.so")
            p = ctypes.c_void_p()
            assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(4 | 1)) == 0      # cudaHostAllocWriteCombined | Portable
            hp = p.value
        import ctypes as C
        rt = C.CDLL("libcudart.so")
        rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        stream = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            assert rt.cudaMemcpyAsync(hp, d.data_ptr(), nbytes, 2, stream) == 0
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n, t0 = 0, time.perf_counter()
        e0.record()
        while time.perf_counter() - t0 < args.seconds:
            assert rt.cudaMemcpyAsync(hp, d.data_ptr(), nbytes, 2, stream) == 0
            n += 1
            if n % 4 == 0:
                torch.cuda.current_stream().synchronize()
        e1.record(); torch.cuda.synchronize()
        gbs = n * nbytes / (e0.elapsed_time(e1) / 1e3) / 1e9
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        allv = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(allv, t)
        else:
            allv = [t]
        res[kind] = {"per_rank_GBps": [float(v.item()) for v in allv], "aggregate_GBps": float(sum(v.item() for v in allv))}
    if rank == 0:
        print(json.dumps({"n_gpus": world, "copy_mb": args.mb, "seconds": args.seconds, "d2h": res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
