"""Device -> host hand-off of results for the reference-facing calls (`as_numpy=True`).

The reference returns numpy arrays; `tensor.cpu().numpy()` would stage every result through pageable memory
(a synchronous, unpinned copy per array).  Here all results of one call are copied asynchronously into PINNED host
buffers and the stream is synchronised once.  The buffers come from torch's caching host allocator: a returned array owns
its buffer (nothing a later call does can overwrite it), and once the caller drops the array the pinned block is recycled,
so a loop that consumes its results pays the page-locking cost only on the first iterations."""
from . import engine


def to_host(tensors):
    """tensors: CUDA tensors (dense, possibly permuted views) or None -> numpy arrays (same shapes and strides) / None."""
    t = engine.torch()
    hosts = []
    for a in tensors:
        if a is None:
            hosts.append(None)
            continue
        if not a.is_cuda:
            hosts.append(a)
            continue
        h = t.empty_strided(tuple(a.shape), tuple(a.stride()), dtype=a.dtype, pin_memory=True) if a.numel() else t.empty_like(a, device="cpu")
        h.copy_(a, non_blocking=True)
        hosts.append(h)
    t.cuda.current_stream().synchronize()
    return [None if h is None else h.numpy() for h in hosts]
