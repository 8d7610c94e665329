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


SUMMARY_SCHEMA = None   # filled by summary_schema(ld)


def summary_schema(ld, ncounters=16):
    """Per-chain outputs of Cells.mcmc_run that the gather assembles: name -> (trailing shape, dtype)."""
    return {"mean": ((ld,), np.float64), "std": ((ld,), np.float64), "sig": ((2,), np.float64),
            "counters": ((ncounters,), np.int64)}


def fit_sharded(run_local, chain_cell, N_of_cell, arrays, chain_uid, rank, world, group=None, schema=None):
    """Run `run_local(chain_cell_slice, arrays_slice, uid_slice) -> dict of per-chain arrays` on this rank's contiguous
    slice of the chains (partition by cumulative work, `parfor` over `numParPools` workers) and gather the per-chain
    outputs so that every rank ends up with the full result in the original chain order.

    The gather is the path's only communication: ONE all_gather per dtype of the packed summaries (float64: mean | std |
    sig, int64: counters), padded to the largest slice — tensors on the GPU over NCCL (NVLink/NVSwitch), on the host over
    gloo.  `schema` (name -> (trailing shape, dtype), e.g. summary_schema(ld)) names the arrays to gather; by default every
    array of the local result whose first dimension is the slice length.  Other entries of the local dict stay local
    (returned under "local")."""
    parts = partition(chain_work(np.asarray(N_of_cell)[chain_cell]), world)
    s, e = parts[rank]
    n_loc = e - s
    local = run_local(chain_cell[s:e], [a[s:e] for a in arrays], chain_uid[s:e]) if n_loc > 0 else {}
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    if schema is None:
        schema = {k: (tuple(np.asarray(v).shape[1:]), np.asarray(v).dtype) for k, v in local.items()
                  if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == n_loc}
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    n_max = max(b - a for a, b in parts)
    out = {}
    for dt in sorted({np.dtype(d).str for _, d in schema.values()}):
        keys = [k for k, (_, d) in schema.items() if np.dtype(d).str == dt]
        widths = [int(np.prod(schema[k][0])) if schema[k][0] else 1 for k in keys]
        pack = np.zeros((n_max, sum(widths)), dtype=np.dtype(dt))
        o = 0
        for k, w in zip(keys, widths):
            if n_loc > 0:
                pack[:n_loc, o:o + w] = np.asarray(local[k]).reshape(n_loc, w)
            o += w
        send = torch.from_numpy(pack).to(dev)
        recv = torch.empty((world * n_max, send.shape[1]), dtype=send.dtype, device=dev)
        dist.all_gather_into_tensor(recv, send, group=group)
        full = recv.cpu().numpy().reshape(world, n_max, send.shape[1])
        o = 0
        for k, w in zip(keys, widths):
            out[k] = np.concatenate([full[r, :b - a, o:o + w] for r, (a, b) in enumerate(parts)], axis=0).reshape(
                (len(chain_cell),) + tuple(schema[k][0]))
            o += w
    out["local"] = {k: v for k, v in local.items() if k not in schema}
    out["partition"] = parts
    return out
