"""Multi-GPU plumbing: one process per GPU, replicas sharded by contiguous global id ranges.

Replicas never exchange state, so the data path has no collective.  The only exchange is at
the end of a protocol: the 8-double moment vector (count, sums, cross sums) is all-reduced
with `torch.distributed` (NCCL over NVLink on GPUs, gloo in CPU tests) and every rank then
evaluates estimate_nonequilibrium_free_energy on the global vector
(reference: benchmark/evaluation/analysis.py:8-17).  Because Philox streams are keyed by the
GLOBAL replica id, the per-replica works are bit-identical for any number of ranks.
"""
import os


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process if absent)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)),
            int(os.environ.get("LOCAL_RANK", 0)))


def shard_range(n_total, rank, world_size):
    """Contiguous shard [lo, hi) of `n_total` replicas owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_process_group(backend=None):
    import torch
    import torch.distributed as dist
    rank, world_size, local_rank = world()
    if world_size > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    from . import _rngstate
    _rngstate.set_rank(rank)  # disjoint replica-id (Philox stream) ranges per rank
    if os.environ.get("LSI_BIND_NUMA", "1") != "0":
        bind_to_gpu_cpus(local_rank)
    return rank, world_size, local_rank


def bind_to_gpu_cpus(local_rank):
    """Pin this process to the CPU cores NVML reports as local to its GPU (same NUMA node / PCIe root), so that pinned
    staging buffers are allocated and the D2H copies of the works are driven from memory next to the device.  Returns the
    cpu set, or None when NVML or the affinity call is unavailable (never fatal)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:  # noqa: BLE001
        pass
    return None


def all_reduce_moments(moments):
    """Sum moment vectors (or a stack of them, shape (..., 8)) over ranks, in place."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(moments, op=dist.ReduceOp.SUM)
    return moments


def all_reduce_moments_async(moments):
    """The same all-reduce started on NCCL's own stream: the caller's stream does not wait for it.  Returns a handle
    whose .wait() orders the current stream after the reduction (None when there is one rank).  Lets the exchange of
    one protocol's moment vectors overlap the next protocol's kernel instead of sitting between two launches."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.all_reduce(moments, op=dist.ReduceOp.SUM, async_op=True)
    return None


def logmeanexp_partials(w):
    """This rank's partial of rows of work samples (one row per outer sample; a 2-D CUDA tensor, or a list of 1-D
    tensors / arrays): (row max of -w, sum exp(-w - max), count), computed by lsi_reduce_logmeanexp on the device
    (log-mean-exp = log(sum / count) + max  =>  sum = count exp(estimate - max))."""
    import torch
    from . import engine
    flat, offs = engine._ragged(w)
    est, _ = engine.reduce_logmeanexp((flat, offs))
    counts = (offs[1:] - offs[:-1]).to(torch.float64)
    seg = torch.repeat_interleave(torch.arange(len(counts), device=flat.device), (offs[1:] - offs[:-1]))
    row_max = torch.full((len(counts),), -float("inf"), dtype=torch.float64, device=flat.device).scatter_reduce(
        0, seg, -flat, reduce="amax", include_self=True)
    row_sumexp = counts * torch.exp(est - row_max)
    return row_max, row_sumexp, counts


def all_reduce_logmeanexp(row_max, row_sumexp, row_count):
    """Merge per-rank partial (max, sum exp(-w - max), count) of the rows of a nested estimator
    (reference: compare_near_eq_and_exact.py:95-97) into the global log-mean-exp per row.  The partials of a shard
    come from logmeanexp_partials (lsi_reduce_logmeanexp); used when the inner samples of ONE outer sample are
    spread over ranks (the sharded outer loop keeps a row on one rank and does not need it)."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        gmax = row_max.clone()
        dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
        scaled = row_sumexp * torch.exp(row_max - gmax)
        scaled = torch.where(torch.isfinite(row_max), scaled, torch.zeros_like(scaled))
        dist.all_reduce(scaled, op=dist.ReduceOp.SUM)
        cnt = row_count.clone()
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        return torch.log(scaled / cnt) + gmax
    return torch.log(row_sumexp / row_count) + row_max
