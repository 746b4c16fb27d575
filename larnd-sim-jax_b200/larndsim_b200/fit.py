"""Fit / scan steps on top of the hot path: what optimize/fit_params.py:704-735 does per iteration in the reference
(``jax.value_and_grad(params_loss)`` on one batch, then an optimiser update) and what its likelihood scans do per grid
point (:1140-1149), with events sharded over the ranks of a torchrun launch.

The loss is the reference's ``mse_adc`` (losses_jax.py:58-82): a charge-weighted MMD between simulated and target hits
plus a total-charge term.  Both are normalised by GLOBAL sums, so a sharded evaluation is two-phase:
every rank forms its local kernel sums, the seven sums are all-reduced (differentiably: the upstream gradient of the
global sums is the local one), every rank back-propagates through its own events and the parameter gradients are
all-reduced.  Two tiny collectives per step; no data-path exchange.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import parallel, sim
from .losses import adc2charge, mse_adc


class FitProblem:
    """One batch of events on this rank + the target hits they are fitted to.

    names        fitted Params fields (leaves of build_params_class(names))
    params       a Params object of that class holding the static configuration
    tracks       (N, n_fields) CUDA tensor of THIS rank's events, local event ids 0 .. n_events-1
    target       the 8-tuple simulate_stochastic returned for the target parameters on the same events
    group        process group the loss sums / gradients are reduced over (None: the default group when initialised)
    """

    def __init__(self, names, params, response_template, tracks, fields, n_events, target, group=None, sigma=1.0, lambda_Q=1.0,
                 distributed=None):
        self.names = tuple(names)
        self.params, self.bank, self.tracks, self.fields, self.n_events = params, response_template, tracks, tuple(fields), int(n_events)
        self.group = group
        self.distributed = (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1) if distributed is None \
            else bool(distributed)
        self.sigma, self.lambda_Q = sigma, lambda_Q
        t = [x.detach() for x in target]
        self.ref = (adc2charge(t[0], params), t[1], t[2], t[3], t[4], t[5], t[6].float())
        self.npix_capacity = None   # fixed after the first evaluation: later steps run without the sizing read-back

    @classmethod
    def from_target_params(cls, names, params, target_values, response_template, tracks, fields, n_events, **kw):
        """Target hits = the same events simulated with ``target_values`` (the closure test every reference fit script runs)."""
        p_tgt = params.replace(**{k: float(v) for k, v in target_values.items()})
        with torch.no_grad():
            w, u = sim.simulate_wfs(p_tgt, response_template, tracks, fields, n_events=n_events)
            target = [x.clone() for x in sim.simulate_stochastic(p_tgt, w, u, 0)]
        return cls(names, params, response_template, tracks, fields, n_events, target, **kw)

    def _reduce(self, x):
        if not self.distributed:
            return x
        return _GroupSum.apply(x, self.group)

    def loss(self, values):
        """Global mse_adc loss for ``values`` (name -> float or 0-d tensor; tensors with requires_grad get gradients)."""
        p = self.params.replace(**values)
        if self.npix_capacity is None:
            wfs, upix = sim.simulate_wfs(p, self.bank, self.tracks, self.fields, n_events=self.n_events)
            # head-room for the pixel list: parameters move during a fit and with them (slightly) the set of main pixels
            self.npix_capacity = sim.pad_size(int(upix.shape[0] * 1.1) + 8, "fit_unique_pixels", 0.2)
        wfs, upix = sim.simulate_wfs(p, self.bank, self.tracks, self.fields, npix_capacity=self.npix_capacity, n_events=self.n_events)
        adcs, x, y, z, ticks, hp, ev, _ = sim.simulate_stochastic(p, wfs, upix, 0)
        loss, aux = mse_adc(p, adc2charge(adcs, p), x, y, z, ticks, hp, ev.float(), *self.ref, sigma=self.sigma,
                            lambda_Q=self.lambda_Q, reduce=self._reduce if self.distributed else None)
        return loss, aux

    def loss_and_grads(self, values):
        """(loss, d loss / d names) for plain-float ``values``: one forward + backward, gradients summed over the ranks."""
        leaves = {n: torch.tensor(float(values[n]), dtype=torch.float32, requires_grad=True) for n in self.names}
        other = {k: v for k, v in values.items() if k not in leaves}
        loss, _ = self.loss(dict(other, **leaves))
        loss.backward()
        dev = self.tracks.device
        g = torch.stack([leaves[n].grad.to(dev) if leaves[n].grad is not None else torch.zeros((), device=dev) for n in self.names])
        if self.distributed:
            dist.all_reduce(g, group=self.group)
        return loss.detach(), g


class _GroupSum(torch.autograd.Function):
    """y = sum over the ranks of ``group`` of x; every rank evaluates the same global loss from y, so the gradient flowing
    back to the local x is the upstream gradient itself (parallel._AllReduceSum with an explicit group)."""

    @staticmethod
    def forward(ctx, x, group):
        y = x.clone()
        dist.all_reduce(y, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


class AdamFit:
    """The reference's fit loop body (optimize/fit_params.py:731-760): parameters normalised by their nominal values,
    Adam on the normalised vector, one loss + gradient evaluation per step."""

    def __init__(self, problem, nominal, lr=0.01):
        self.problem, self.nominal = problem, {n: float(nominal[n]) for n in problem.names}
        dev = problem.tracks.device
        self.theta = torch.ones(len(problem.names), device=dev, requires_grad=True)
        self.opt = torch.optim.Adam([self.theta], lr=lr)
        self._scale = torch.tensor([self.nominal[n] for n in problem.names], device=dev)

    def step(self):
        self.opt.zero_grad(set_to_none=False)
        vals = {n: self.theta[i] * self.nominal[n] for i, n in enumerate(self.problem.names)}
        loss, _ = self.problem.loss(vals)
        loss.backward()
        if self.problem.distributed:
            dist.all_reduce(self.theta.grad, group=self.problem.group)
        self.opt.step()
        return loss.detach()

    def values(self):
        return {n: float(v) for n, v in zip(self.problem.names, (self.theta.detach() * self._scale).cpu())}


def scan_layout(world_size, event_shards):
    """Ranks as a (point groups) x (event shards) grid: rank = point_group * event_shards + event_shard."""
    if world_size % event_shards:
        raise ValueError("world size %d is not a multiple of %d event shards" % (world_size, event_shards))
    return world_size // event_shards, event_shards


_group_cache = {}


def make_shard_groups(world_size, event_shards):
    """Process groups of the ranks that share a grid point (one per point group).  Collective: call on every rank.  Cached:
    communicators are created once per (world size, shard count)."""
    key = (world_size, event_shards)
    if key not in _group_cache:
        groups = []
        for pg in range(world_size // event_shards):
            ranks = list(range(pg * event_shards, (pg + 1) * event_shards))
            groups.append(dist.new_group(ranks) if event_shards > 1 else None)
        _group_cache[key] = groups
    return _group_cache[key]


def scan_2d(make_problem, p1, a1, p2, a2, fixed=None, event_shards=1):
    """Loss and gradients on the grid a1 x a2 over (p1, p2) — BASELINE config 5's "2-D likelihood scan".  Grid points are
    dealt round-robin to the point groups, the events of a point are sharded over the ``event_shards`` ranks of its group
    (make_problem(event_shard, event_shards, group) -> FitProblem with names (p1, p2)), and one all-gather collects the
    (loss, dloss/dp1, dloss/dp2) table.  Returns an (len(a1), len(a2), 3) numpy array (identical on every rank)."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    n_pg, n_es = scan_layout(world, event_shards)
    pg, es = divmod(rank, n_es)
    groups = make_shard_groups(world, n_es) if world > 1 else [None]
    prob = make_problem(es, n_es, groups[pg])
    prob.distributed = n_es > 1   # ranks of DIFFERENT point groups evaluate different points: nothing is reduced across them
    n1, n2 = len(a1), len(a2)
    mine = list(range(pg, n1 * n2, n_pg))
    per_group = (n1 * n2 + n_pg - 1) // n_pg
    local = torch.zeros((per_group, 4), device=prob.tracks.device)
    local[:, 0] = -1
    for k, ipt in enumerate(mine):
        i, j = divmod(ipt, n2)
        loss, g = prob.loss_and_grads(dict(fixed or {}, **{p1: float(a1[i]), p2: float(a2[j])}))
        local[k, 0] = float(ipt)
        local[k, 1] = loss
        local[k, 2:4] = g
    table = parallel.allgather(local).reshape(-1, 4).cpu().numpy()
    out = np.full((n1, n2, 3), np.nan)
    for ipt, l, g1, g2 in table:
        if ipt >= 0:
            out[int(ipt) // n2, int(ipt) % n2] = (l, g1, g2)
    return out
