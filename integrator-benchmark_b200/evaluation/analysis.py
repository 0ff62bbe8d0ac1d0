"""Drop-in for benchmark/evaluation/analysis.py:8-17."""
from .. import engine


def estimate_nonequilibrium_free_energy(W_shads_F, W_shads_R, N_eff=None):
    """DeltaF_neq = (mean W_F - mean W_R) / 2 and its squared uncertainty
    [var(W_F) + var(W_R) - 2 cov(W_F, W_R)] / (4 N_eff), with np.var's ddof = 0 and np.cov's
    ddof = 1 exactly as the reference mixes them.  The six sums are reduced on the GPU
    (lsi_reduce_moments); inputs may be numpy arrays or CUDA tensors."""
    m = engine.reduce_moments(W_shads_F, W_shads_R).cpu().numpy()
    n = len(W_shads_F)
    if m[6] > 0:  # non-finite works poison np.mean in the reference as well
        return float("nan"), float("nan")
    return engine.neq_free_energy_from_moments(m, N_eff if N_eff is not None else n)


def estimate_from_moments(moments, N_eff=None):
    """Same estimate from an (all-reduced) moment vector; non-finite pairs are left out."""
    return engine.neq_free_energy_from_moments([float(v) for v in moments], N_eff)
