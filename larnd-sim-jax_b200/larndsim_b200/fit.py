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


def scan_2d(make_problem, p1, a1, p2, a2, fixed=None, event_shards=1, fused=True):
    """Loss and gradients on the grid a1 x a2 over (p1, p2) — BASELINE config 5's "2-D likelihood scan".  Grid points are
    dealt round-robin to the point groups, the events of a point are sharded over the ``event_shards`` ranks of its group
    (make_problem(event_shard, event_shards, group) -> FitProblem with names (p1, p2)), and one all-gather collects the
    (loss, dloss/dp1, dloss/dp2) table.  Returns an (len(a1), len(a2), 3) numpy array (identical on every rank).
    ``fused`` (default): every point is one FusedFitStep call (one asynchronous chain of C-ABI calls, ~0.9 ms at 20 k segments)
    instead of FitProblem.loss_and_grads under torch.autograd (~3.4 ms)."""
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
    step = FusedFitStep(prob) if fused else None
    rows = []
    for k, ipt in enumerate(mine):
        i, j = divmod(ipt, n2)
        vals = dict(fixed or {}, **{p1: float(a1[i]), p2: float(a2[j])})
        if fused:
            loss, g = step(vals)
            rows.append((float(ipt), float(loss), float(g[0]), float(g[1])))
        else:
            loss, g = prob.loss_and_grads(vals)
            local[k, 0] = float(ipt)
            local[k, 1] = loss
            local[k, 2:4] = g
    if fused and rows:
        local[:len(rows)] = torch.tensor(rows, dtype=local.dtype, device=local.device)
    table = parallel.allgather(local).reshape(-1, 4).cpu().numpy()
    out = np.full((n1, n2, 3), np.nan)
    for ipt, l, g1, g2 in table:
        if ipt >= 0:
            out[int(ipt) // n2, int(ipt) % n2] = (l, g1, g2)
    return out


class FusedFitStep:
    """The same loss + gradients as FitProblem.loss_and_grads, as ONE asynchronous chain of kernel launches through the C ABI:
    prepare -> unique -> accumulate -> front end (dense, no hit compaction) -> dense mse_adc sums -> [all-reduce 5 floats]
    -> mse_adc VJP -> front-end VJP (step events) -> accumulate VJP + chain rule -> [all-reduce 15 floats] -> one read-back.
    Every buffer is allocated once; no autograd graph, no torch glue kernels, a single host synchronisation per step (the
    read of loss + gradients the optimiser needs anyway).  This is the regime every optimize/ fit of the reference runs in
    (~20 k segments per batch, optimize/fit_test.sh), where a step is bound by host work, not by the kernels."""

    def __init__(self, problem, capacity_margin=1.1):
        import ctypes as C
        from . import _lib
        self.C, self._lib, self.lib = C, _lib, _lib.get_lib()
        pr = self.problem = problem
        dev = self.dev = pr.tracks.device
        lib = self.lib
        # static configuration: a Params object WITHOUT tensor leaves (values are plain floats here)
        self.params = pr.params
        with torch.cuda.device(dev):
            self.lut = sim.get_lut(pr.bank, pr.params.signal_length, pr.params.nb_sampling_bins_per_pixel, pr.params.number_pix_neighbors)
            self.pod = sim.make_pod(self._with_values({}), self.lut.shape)
            self.cols = sim.make_columns(pr.fields)
            n = self.n = pr.tracks.shape[0]
            self.tracks = pr.tracks.contiguous()
            self.ws_bytes = lib.larnd_workspace_bytes(n, pr.n_events, self.pod.n_tpc, self.pod.n_pixels_x, self.pod.n_pixels_y)
            self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
            self.counts = torch.zeros(4, dtype=torch.int32, device=dev)
            st = sim.lut_forward(self._with_values({}), pr.bank, self.tracks, pr.fields, n_events=pr.n_events)
            self.npix = sim.pad_size(int(st.npix * capacity_margin) + 8, "fused_fit_unique_pixels", 0.2)
            del st
            nt, k = self.pod.n_ticks, self.pod.max_adc_values
            npix = self.npix
            self.upix = torch.empty(npix, dtype=torch.int32, device=dev)
            self.wfs = sim.alloc_wfs(npix, nt, dev)
            f32 = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
            self.adc, self.ticks, self.pz = f32(npix, k), f32(npix, k), f32(npix, k)
            self.px, self.py = f32(npix), f32(npix)
            self.event = torch.empty(npix, dtype=torch.int32, device=dev)
            self.saved = torch.zeros((npix, 32), dtype=torch.float32, device=dev)
            self.n_valid = torch.zeros(1, dtype=torch.int32, device=dev)
            self.fee_scratch = torch.empty(lib.larnd_fee_scratch_bytes(npix), dtype=torch.uint8, device=dev)
            self.g_adc = f32(npix, k)
            # backward: step events for batches the class-sorted tile kernels serve (>= 200 k segments); below that the chunk
            # kernel is faster on the dense gradient rows (measured: 1.5 vs 1.9 ms per 19.8 k-segment step)
            self.use_steps = self.n >= 200_000
            if self.use_steps:
                self.steps = torch.empty(lib.larnd_fee_steps_bytes(npix), dtype=torch.uint8, device=dev)
            else:
                self.g_wfs = torch.zeros((npix, self.wfs.shape[1]), dtype=torch.float32, device=dev)   # padding / column 0 stay zero
            # target side of the loss: points, weights, Kyy, Sy (constants of the fit)
            rq, rx, ry, rz, _rt, rhp, rev = pr.ref
            self.ref_pts = torch.stack((rx + rev * 1e5, ry, rz), dim=-1).contiguous().float()
            self.ref_w = (rq * rhp).contiguous().float()
            self.n_ref = int(self.ref_w.shape[0])
            from .losses import rbf_field
            fyy = rbf_field(self.ref_pts, self.ref_pts, self.ref_w, pr.sigma) if self.n_ref else torch.zeros((0, 4), device=dev)
            self.out = torch.zeros(8 + 4 + _lib.NPARAMS + 4, dtype=torch.float32, device=dev)   # sums | loss | grads | counts (as float)
            self.sums, self.loss, self.grad = self.out[0:8], self.out[8:12], self.out[12:12 + _lib.NPARAMS]
            self.kyy_sy = torch.stack([(self.ref_w * fyy[:, 0]).sum(), self.ref_w.sum()]) if self.n_ref else torch.zeros(2, device=dev)
            self.loss_scratch = torch.empty(max(lib.larnd_mse_adc_scratch_bytes(npix, k, self.n_ref), 256), dtype=torch.uint8, device=dev)
            self.host = torch.zeros(self.out.shape[0], dtype=torch.float32).pin_memory()
        self.idx = [_lib.PARAM_ORDER.index(nm) for nm in pr.names]

    def _with_values(self, values):
        """A plain-float Params object (same fitted-field list, no tensor leaves) carrying ``values``."""
        from .consts import _DEFAULTS, float_params_class
        if getattr(self, "_float_base", None) is None:
            pp = self.problem.params
            d = {k: getattr(pp, k) for k in _DEFAULTS}
            for nm in self.problem.names:
                d[nm] = float(pp.value(nm))
            for k, v in d.items():
                if torch.is_tensor(v) and v.numel() == 1:
                    d[k] = float(pp.value(k))
            self._float_cls, self._float_base = float_params_class(type(pp)), d
        return self._float_cls(**dict(self._float_base, **{nm: float(v) for nm, v in values.items()}))

    def __call__(self, values):
        """(loss, gradients w.r.t. problem.names as a numpy array) for plain-float parameter values."""
        C, lib, pr = self.C, self.lib, self.problem
        ptr = lambda t: C.c_void_p(t.data_ptr())
        check = self._lib.check
        with torch.cuda.device(self.dev):
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            pod = sim.fill_pod_leaves(self.pod, self._with_values(values))
            P = C.byref(pod)
            npix, nt = self.npix, pod.n_ticks
            self.out[:12 + self._lib.NPARAMS].zero_()
            self.sums[3:5].copy_(self.kyy_sy)
            self.counts.zero_()
            check(lib.larnd_lut_forward(ptr(self.tracks), self.n, C.byref(self.cols), P, self.lut.handle, pr.n_events, npix, 0,
                                        ptr(self.workspace), self.ws_bytes, ptr(self.upix), ptr(self.wfs), self.wfs.stride(0),
                                        ptr(self.counts), stream))
            wfs1 = C.c_void_p(self.wfs.data_ptr() + 4)          # simulate_wfs' view [:, 1:]
            null = C.c_void_p(0)
            check(lib.larnd_fee_forward(wfs1, self.wfs.stride(0), ptr(self.upix), npix, P, null, ptr(self.adc), ptr(self.ticks),
                                        ptr(self.pz), ptr(self.px), ptr(self.py), ptr(self.event), ptr(self.saved),
                                        null, null, null, null, null, null, null, null, ptr(self.n_valid),
                                        ptr(self.fee_scratch), self.fee_scratch.numel(), stream))
            fee_out = (ptr(self.adc), ptr(self.ticks), ptr(self.pz), ptr(self.px), ptr(self.py), ptr(self.event), ptr(self.upix))
            check(lib.larnd_mse_adc_sums(*fee_out, npix, P, ptr(self.ref_pts), ptr(self.ref_w), self.n_ref, float(pr.sigma),
                                         ptr(self.sums), ptr(self.loss_scratch), self.loss_scratch.numel(), stream))
            if pr.distributed:
                dist.all_reduce(self.sums[:5], group=pr.group)
            check(lib.larnd_mse_adc_backward(ptr(self.sums), *fee_out, npix, P, self.n_ref, float(pr.sigma), float(pr.lambda_Q),
                                             ptr(self.loss), ptr(self.g_adc), ptr(self.grad), ptr(self.loss_scratch),
                                             self.loss_scratch.numel(), stream))
            if self.use_steps:   # the two VJPs without the dense (npix, n_ticks) waveform gradient: step events per pixel row
                check(lib.larnd_fee_backward_steps(ptr(self.g_adc), ptr(self.saved), ptr(self.upix), npix, P, ptr(self.steps),
                                                   self.steps.numel(), 0, stream))
                check(lib.larnd_lut_backward_steps(self.n, P, self.lut.handle, pr.n_events, npix, self._lib.FLAG_REUSE_RUNS, ptr(self.workspace), self.ws_bytes,
                                                   ptr(self.counts), ptr(self.steps), self.steps.numel(), ptr(self.grad), stream))
            else:
                g1 = C.c_void_p(self.g_wfs.data_ptr() + 4)
                check(lib.larnd_fee_backward(ptr(self.g_adc), ptr(self.ticks), ptr(self.saved), npix, P, g1, self.g_wfs.stride(0), 0, stream))
                check(lib.larnd_lut_backward(self.n, P, self.lut.handle, pr.n_events, npix, 1 | self._lib.FLAG_REUSE_RUNS, ptr(self.workspace), self.ws_bytes,
                                             ptr(self.counts), ptr(self.g_wfs), self.g_wfs.stride(0), ptr(self.grad), stream))
            if pr.distributed:
                dist.all_reduce(self.grad, group=pr.group)
            self.out[12 + self._lib.NPARAMS:].copy_(self.counts)   # device-side flags travel with the results
            self.host.copy_(self.out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        h = self.host.numpy()
        flags = int(h[12 + self._lib.NPARAMS + 2])
        if flags:
            raise self._lib.LarndError("fit step: device-side check failed (flags %d: 1 = pixel capacity %d too small, 2 = event id "
                                       "out of range, 4 = template row beyond the bank)" % (flags, self.npix))
        return float(h[8]), h[12:12 + self._lib.NPARAMS][self.idx].astype(np.float64)


class FusedAdamFit:
    """AdamFit with FusedFitStep: parameters live on the host (six floats), Adam runs on the host, one synchronisation per
    step.  Same update rule as torch.optim.Adam (bias-corrected first / second moments, eps outside the square root)."""

    def __init__(self, problem, nominal, lr=0.01, betas=(0.9, 0.999), eps=1e-8):
        self.step_fn = FusedFitStep(problem)
        self.names = problem.names
        self.nominal = np.array([float(nominal[n]) for n in self.names])
        self.theta = np.ones(len(self.names))
        self.lr, self.b1, self.b2, self.eps = lr, betas[0], betas[1], eps
        self.m, self.v, self.t = np.zeros_like(self.theta), np.zeros_like(self.theta), 0

    def step(self):
        vals = {n: float(t * s) for n, t, s in zip(self.names, self.theta, self.nominal)}
        loss, g = self.step_fn(vals)
        g = g * self.nominal                       # d loss / d theta (theta = value / nominal)
        self.t += 1
        self.m = self.b1 * self.m + (1 - self.b1) * g
        self.v = self.b2 * self.v + (1 - self.b2) * g * g
        mh, vh = self.m / (1 - self.b1 ** self.t), self.v / (1 - self.b2 ** self.t)
        self.theta = self.theta - self.lr * mh / (np.sqrt(vh) + self.eps)
        return loss

    def values(self):
        return {n: float(t * s) for n, t, s in zip(self.names, self.theta, self.nominal)}


FastAdamFit = FusedAdamFit
