"""Multi-GPU plumbing: one process per GPU, events sharded across ranks, tiny collectives only.

The reference is single-device.  The unit of work that shards is the *event*: the event id is part
of the pixel key (detsim_jax.py:241-242), so two events never share a waveform row and every rank can run the whole
prepare -> accumulate -> FEE pipeline on its own events with rank-local event ids.  The only exchange steps are
  * a sum all-reduce of [loss terms..., d loss / d theta (<= 15 floats)] per fit step, and
  * an all-gather of (loss, gradients) per grid point for likelihood scans.
Both are O(100) bytes: NCCL over NVLink when the tensors live on CUDA devices, gloo on CPU (used by the tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index, sysfs="/sys/bus/pci/devices"):
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off, BEFORE it allocates pinned host buffers
    (first touch places them on that node).  With one process per GPU, every rank's host-to-device stream then stays on its
    own socket's memory controllers and PCIe root instead of crossing the inter-socket link.  Returns the CPU set used, or
    None when the topology is unknown (no sysfs entry, a single node, or a cpuset that excludes those CPUs) — never raises."""
    import os
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(os.path.join(sysfs, bdf, "local_cpulist")) as fh:
            local = _parse_cpulist(fh.read())
        allowed = os.sched_getaffinity(0)
        cpus = local & allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def event_partition(event_ids, world_size):
    """Contiguous event ranges balanced by segment count.  ``event_ids``: per-row local event id (>= 0; padding rows
    with -1 are ignored).  Returns a list of (first_event, last_event_exclusive) per rank."""
    ev = np.asarray(event_ids).astype(np.int64)
    ev = ev[ev >= 0]
    n_events = int(ev.max()) + 1 if ev.size else 0
    counts = np.bincount(ev, minlength=n_events).astype(np.int64)
    cum = np.concatenate([[0], np.cumsum(counts)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        e = int(np.searchsorted(cum, target, side="left"))
        bounds.append(min(max(e, bounds[-1]), n_events))
    bounds.append(n_events)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def shard_tracks(tracks, fields, rank, world_size):
    """Rows of this rank's events with event ids renumbered to start at 0 (batch-local ids, as the reference's
    remap_event_ids_to_local does per batch, optimize/dataio.py:47-61).  Works on numpy arrays and torch tensors.
    Returns (local_tracks, n_local_events, first_global_event)."""
    col = tuple(fields).index("eventID")
    is_t = torch.is_tensor(tracks)
    ev = tracks[:, col].detach().cpu().numpy() if is_t else np.asarray(tracks[:, col])
    lo, hi = event_partition(ev, world_size)[rank]
    sel = (ev >= lo) & (ev < hi)
    if is_t:
        local = tracks[torch.as_tensor(sel, device=tracks.device)].clone()
    else:
        local = np.array(tracks[sel], copy=True)
    local[:, col] -= lo
    return local, hi - lo, lo


def allreduce_sum_(t):
    """In-place sum over ranks (no-op without an initialised process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allgather(t):
    """Gathers equally-shaped tensors from every rank -> (world, *shape)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t.unsqueeze(0)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out)


def scan_points_for_rank(n_points, rank, world_size):
    """Round-robin assignment of likelihood-scan grid points to ranks."""
    return list(range(rank, n_points, world_size))


class _AllReduceSum(torch.autograd.Function):
    """y = sum over ranks of x.  Every rank evaluates the same global loss from y, so the gradient that flows back to the
    local x is the upstream gradient itself; the parameter gradients of the ranks are then summed by the caller
    (allreduce_sum_) — the two-phase scheme of SURVEY.md §8e for losses normalised by global sums."""

    @staticmethod
    def forward(ctx, x):
        y = x.clone()
        allreduce_sum_(y)
        return y

    @staticmethod
    def backward(ctx, g):
        return g


def allreduce_sum_differentiable(x):
    return _AllReduceSum.apply(x)
