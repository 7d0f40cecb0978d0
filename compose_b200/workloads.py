"""Synthetic cubed-sphere workloads for CDR::run.

A BASELINE.json config "ne<N> x <nlev> levels x <q> tracers" is realised as
ncells = 6*N^2 leaf cells and nt = nlev*q independent CDR tracers over the
recursive-bisection tree of the space-filling-curve cell order
(tree::make_tree_over_1d_mesh). Inputs are a deterministic splitmix64 stream so
that the host (numpy, here) and the device (cedr_b200_fill_headline) generators
produce bit-identical arrays.
"""
import numpy as np

CONFIGS = {
    # name: (ncells, nt, config_id)
    "ne30x72x40": (6*30*30, 72*40, 1),
    "ne120x128x40": (6*120*120, 128*40, 2),
    "ne256x128x10": (6*256*256, 128*10, 3),
}

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix_u(seed, k):
    """U(k) = (splitmix64 output k of stream `seed`) >> 11, times 2^-53."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (np.asarray(k, dtype=np.uint64) + np.uint64(1))*_GOLDEN
        z = (z ^ (z >> np.uint64(30)))*_M1
        z = (z ^ (z >> np.uint64(27)))*_M2
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64)*2.0**-53


def headline(ncells, nt, config_id, t0=0, t1=None):
    """rhom[ncells] and qm_min, qm, qm_max, qm_prev [t1-t0, ncells] for tracers
    [t0, t1): rhom = 0.5(1+U); q_min = 0.1U; q_max = q_min + U;
    q = q_min + (q_max - q_min)(1.4U - 0.2) (about 29% of cells out of bounds);
    q_prev = q_min + (q_max - q_min)U; Qm_* = q_* rhom."""
    t1 = nt if t1 is None else t1
    seed = 0xCED20000 + config_id
    i = np.arange(ncells, dtype=np.uint64)
    rhom = 0.5*(1 + splitmix_u(seed, i))
    k = (np.arange(t0, t1, dtype=np.uint64)[:, None]*np.uint64(ncells) + i[None, :])
    p = np.uint64(ncells) + np.uint64(4)*k
    q_min = 0.1*splitmix_u(seed, p)
    q_max = q_min + splitmix_u(seed, p + np.uint64(1))
    q = q_min + (q_max - q_min)*(1.4*splitmix_u(seed, p + np.uint64(2)) - 0.2)
    q_prev = q_min + (q_max - q_min)*splitmix_u(seed, p + np.uint64(3))
    return rhom, q_min*rhom, q*rhom, q_max*rhom, q_prev*rhom
