"""Multi-GPU layout of the benchmark and of I-V sweeps: one context per applied bias, one bias per rank.

The reference has no sweep driver (`applied bias` is one scalar, reference input_file.prm:86); BASELINE.json config 5
places one applied voltage per GPU.  Contexts of different biases share nothing, so there is no data-path
collective: torch.distributed is used only for the barrier and for the max-over-ranks of the device time."""
import os


def bias_for_rank(rank, world_size, v_min=0.0, v_step=0.05):
    """applied bias [V] of a rank: 0, 0.05, 0.10, ... (>= 0: the reference's pattern forbids negative values,
    reference source/ParameterReader.cpp:66-68)"""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return v_min + v_step * rank


def init_distributed(backend=None):
    """(rank, local_rank, world_size, dist or None) from the torchrun environment; no process group for 1 rank"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return rank, local, world, None
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if not dist.is_initialized():
        dist.init_process_group(backend=backend or "nccl", rank=rank, world_size=world)
        if (backend or "nccl") == "nccl":
            # rank 0 prints ONE JSON line on stdout: NCCL writes its "NCCL version ..." banner there when the first
            # communicator is created (NCCL_DEBUG=VERSION / WARN on some boxes), so create it now with fd 1 parked on fd 2
            import sys
            import torch
            sys.stdout.flush()
            saved = os.dup(1)
            try:
                os.dup2(2, 1)
                torch.cuda.set_device(local)
                dist.barrier(device_ids=[local])
                torch.cuda.synchronize()
            finally:
                os.dup2(saved, 1)
                os.close(saved)
    return rank, local, world, dist


def pin_to_gpu_numa_node(local_rank, local_world):
    """Pin this rank's host threads to cores of its GPU's NUMA node (NVML's ideal CPU affinity of the device), and to its
    own share of them when several ranks report the same set (the end-to-end sweep moves 50 MB per step
    and rank through host memory).  Returns the core list, or None when NVML / sched_setaffinity are unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        index = local_rank
        if visible and all(t.strip().isdigit() for t in visible.split(",")):
            index = int(visible.split(",")[local_rank])
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        cores = [c for c in cores if c in os.sched_getaffinity(0)]
        if not cores:
            return None
        share = max(1, len(cores) // max(local_world, 1))
        mine = cores[local_rank * share:(local_rank + 1) * share] or cores
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return None


def barrier(dist, device=None):
    if dist is not None:
        if device is not None:
            dist.barrier(device_ids=[device])
        else:
            dist.barrier()


def max_over_ranks(value, dist, device=None):
    """max of a python float over all ranks (gloo on CPU tensors, nccl on the rank's GPU)"""
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=f"cuda:{device}" if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, dist, device=None):
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=f"cuda:{device}" if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_iv(dist, bias, currents):
    """One point of the I-V curve per rank -> the whole curve on every rank, sorted by bias:
    list of (applied bias [V], electron-transfer current, hole-transfer current) in scaled units
    (SolarCellProblem.interface_currents).  The only communication of a sweep, and not on the data path."""
    point = (float(bias), float(currents[0]), float(currents[1]))
    if dist is None:
        return [point]
    points = [None] * dist.get_world_size()
    dist.all_gather_object(points, point)
    return sorted(points)
