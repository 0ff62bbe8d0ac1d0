"""Process-wide RNG bookkeeping for the shims.

The reference draws from the global numpy / OpenMM generators; here every replica owns
Philox streams keyed by (seed, global replica id), so the only global state is the seed and
the next unused replica id.  Two runs with the same seed and the same sequence of calls are
bit-identical, whatever the number of GPUs.
"""
_state = {"seed": 0x5EED0000, "next_replica": 0}


def set_seed(seed, next_replica=0):
    _state["seed"] = int(seed) & 0xFFFFFFFFFFFFFFFF
    _state["next_replica"] = int(next_replica)


def get_seed():
    return _state["seed"]


def reserve_replicas(n):
    """Returns the first of `n` fresh global replica ids."""
    first = _state["next_replica"]
    _state["next_replica"] = first + int(n)
    return first
