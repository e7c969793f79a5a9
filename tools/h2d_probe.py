"""Host->device copy bandwidth of pinned memory, all ranks at once (torchrun): the ceiling of the end-to-end
path when every GPU of a box is fed from the host at the same time.  Prints one JSON line on rank 0."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 1 << 30
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = {}
    for label, together in (("all_ranks_at_once", True), ("one_rank_at_a_time", False)):
        rates = []
        for turn in range(world if not together else 1):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            mine = together or turn == int(os.environ.get("RANK", "0"))
            t0 = time.perf_counter()
            if mine:
                for _ in range(8):
                    d.copy_(host, non_blocking=True)
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                dist.barrier()
            if mine:
                rates.append(8 * nbytes / dt / 1e9)
        r = torch.tensor([rates[0] if rates else 0.0], dtype=torch.float64, device=dev)
        every = [torch.zeros_like(r) for _ in range(world)]
        if world > 1:
            dist.all_gather(every, r)
        else:
            every = [r]
        out[label] = {"per_rank_gbs": [round(float(x.item()), 2) for x in every], "sum_gbs": round(sum(float(x.item()) for x in every), 2)}
    aff = sorted(os.sched_getaffinity(0))
    if int(os.environ.get("RANK", "0")) == 0:
        out["host_cpus"] = len(aff)
        try:
            import pynvml
            pynvml.nvmlInit()
            out["gpu_numa_cpu_affinity_words"] = [list(pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(i), 4)) for i in range(world)]
        except Exception as e:      # noqa: BLE001
            out["nvml"] = repr(e)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
