"""One-process-per-GPU sharding of independent chains (the `parfor (cellNum = 1:N, numParPools)`
of src/TranscriptionCycleMCMC.m:161 across ranks).  The path has NO per-step exchange: chains are
partitioned up front, every rank fits its slice, and one final gather assembles the summaries
(SURVEY.md 8e).  Raw chains, when requested, stay on the rank that produced them."""
import numpy as np


def partition(work, nparts):
    """Contiguous blocks of ~equal total work.  Returns [(start, end)] * nparts (blocks may be empty
    only when there are fewer units than parts)."""
    work = np.asarray(work, dtype=np.float64)
    n = work.size
    cum = np.concatenate([[0.0], np.cumsum(work)])
    bounds = [0]
    for p in range(1, nparts):
        target = cum[-1] * p / nparts
        e = int(np.searchsorted(cum, target, side="left"))
        e = max(e, bounds[-1] + (1 if n - bounds[-1] > nparts - p else 0))
        e = min(e, n)
        bounds.append(e)
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(nparts)]


def chain_work(N_of_chain):
    """Per-step cost model of a chain with N time points: forward model + proposal/covariance
    algebra are both ~quadratic in N."""
    N = np.asarray(N_of_chain, dtype=np.float64)
    return N * N + (N + 7) ** 2


def fit_sharded(run_local, chain_cell, N_of_cell, arrays, chain_uid, rank, world, group=None):
    """Run `run_local(chain_cell_slice, arrays_slice, uid_slice) -> dict of per-chain arrays` on this
    rank's contiguous slice and all-gather the per-chain outputs (small: summaries and counters) so
    that every rank ends up with the full result in the original chain order."""
    import torch.distributed as dist
    parts = partition(chain_work(np.asarray(N_of_cell)[chain_cell]), world)
    s, e = parts[rank]
    local = run_local(chain_cell[s:e], [a[s:e] for a in arrays], chain_uid[s:e]) if e > s else {}
    if world == 1:
        return local
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: np.asarray(v) for k, v in local.items()}, group=group)
    keys = [k for g in gathered for k in g.keys()]
    out = {}
    for k in dict.fromkeys(keys):
        out[k] = np.concatenate([g[k] for g in gathered if k in g], axis=0)
    return out
