"""Process-wide RNG bookkeeping for the shims.

The reference draws from the global numpy / OpenMM generators; here every replica owns
Philox streams keyed by (seed, global replica id), so the only global state is the seed and
the next unused replica id.  Two runs with the same seed and the same sequence of calls are
bit-identical, whatever the number of GPUs.

Multi-process runs (one process per GPU): every rank hands out replica ids from its own range
[rank << 40, (rank + 1) << 40) -- parallel.init_process_group() calls set_rank() -- so that ranks which share
a seed never share a Philox stream (their trajectories would otherwise be identical copies).
"""
_state = {"seed": 0x5EED0000, "next_replica": 0, "rank": 0}
RANK_SHIFT = 40


def set_seed(seed, next_replica=0):
    _state["seed"] = int(seed) & 0xFFFFFFFFFFFFFFFF
    _state["next_replica"] = int(next_replica)


def get_seed():
    return _state["seed"]


def set_rank(rank):
    """Replica ids reserved from now on come from this rank's range."""
    _state["rank"] = int(rank)


def get_rank():
    return _state["rank"]


def reserve_replicas(n):
    """Returns the first of `n` fresh global replica ids."""
    first = _state["next_replica"]
    _state["next_replica"] = first + int(n)
    if _state["next_replica"] >= (1 << RANK_SHIFT):
        raise OverflowError("more than 2^40 replica ids reserved by one process")
    return (_state["rank"] << RANK_SHIFT) + first
