"""CPU-only tests of the host logic: C-ABI surface, parameter container, shape bucketing, synthetic generators,
event sharding with a world_size-2 gloo group.  No compute call is made on the CUDA library."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import common as cm
from oracle import consts as oc
from oracle import larnd_oracle as lo

ROOT = cm.ROOT


@pytest.fixture(scope="module")
def lib():
    import larndsim_b200
    larndsim_b200.build_library()
    return larndsim_b200.get_lib()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "larnd_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(larnd_[a-z_0-9]+)\s*\(", hdr))
    assert {"larnd_lut_forward", "larnd_lut_backward", "larnd_fee_forward", "larnd_fee_backward", "larnd_mc_forward",
            "larnd_mc_backward", "larnd_lut_create", "larnd_workspace_bytes", "larnd_last_error"} <= names
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in include/larnd_b200.h is not exported" % n
    assert lib.larnd_abi_version() == 2


def test_params_pod_layout_matches_the_header(lib, tmp_path):
    """sizeof/offsetof of the ctypes mirror == the C struct compiled from the header."""
    from larndsim_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "larnd_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", '
                   'sizeof(larnd_params_t), offsetof(larnd_params_t, tpc_borders), offsetof(larnd_params_t, long_diff_template), '
                   'offsetof(larnd_params_t, ts_vdrift), sizeof(larnd_columns_t));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    P = _lib.ParamsPOD
    assert got == [ctypes.sizeof(P), P.tpc_borders.offset, P.long_diff_template.offset, P.ts_vdrift.offset, ctypes.sizeof(_lib.Columns)]


def test_workspace_size_and_argument_errors_without_gpu(lib):
    n = lib.larnd_workspace_bytes(1000, 3, 2, 140, 280)
    assert n > 31 * 1000 * 4 and lib.larnd_workspace_bytes(-1, 3, 2, 140, 280) == 0
    assert lib.larnd_lut_create(None, 100, 45, 45, 1950, 100, None, None) != 0
    assert b"invalid" in lib.larnd_last_error()


def test_params_container_mirrors_reference_semantics():
    import torch
    import larndsim_b200 as lb
    P = lb.build_params_class(["Ab", "eField"])
    p = lb.load_geometry_json(P, cm.GEOM)
    assert torch.is_tensor(p.Ab) and p.Ab.requires_grad and not torch.is_tensor(p.kb)
    assert [n for n, _ in p.grad_leaves()] == ["Ab", "eField"]
    q = p.replace(signal_length=77, lifetime=1234.0)
    assert q.signal_length == 77 and p.signal_length == 150 and q.lifetime == 1234.0 and type(q) is type(p)
    with pytest.raises(AttributeError):
        p.signal_length = 3
    with pytest.raises(ValueError):
        lb.build_params_class(["not_a_param"])
    v = lb.get_vdrift(p)
    assert torch.is_tensor(v) and abs(float(v) - oc.get_vdrift(cm.oracle_params())) < 1e-6
    v.backward()
    assert abs(float(p.eField.grad) - 0.1598) < 1e-3          # dv/dE at 0.5 kV/cm
    assert abs(lb.get_vdrift(lb.load_geometry_json(lb.build_params_class([]), cm.GEOM)) - 0.159645) < 1e-6


def test_params_host_snapshot_is_batched_and_tracks_in_place_updates():
    """Params.value fetches every tensor-valued field in one go and keeps the snapshot until a tensor is modified in place
    (an optimiser stepping on a leaf) -- the parameter block is built several times per fit step."""
    import torch
    import larndsim_b200 as lb
    P = lb.build_params_class(["Ab", "kb", "lifetime"])
    p = lb.load_geometry_json(P, cm.GEOM).replace(Ab=torch.tensor(0.75, requires_grad=True))
    assert p.value("Ab") == float(np.float32(0.75)) and p.value("kb") == float(np.float32(p.kb.item()))
    snap = p.__dict__["_host_snapshot"]
    assert set(snap[1]) == {"Ab", "kb", "lifetime"}
    assert p.value("lifetime") == float(p.lifetime.detach()) and p.__dict__["_host_snapshot"] is snap   # served from the snapshot
    with torch.no_grad():
        p.Ab.add_(0.05)                                                                                  # optimiser step in place
    assert abs(p.value("Ab") - 0.8) < 1e-6 and p.__dict__["_host_snapshot"] is not snap
    q = p.replace(kb=0.05)
    assert "_host_snapshot" not in q.__dict__ and abs(q.value("kb") - 0.05) < 1e-8 and abs(q.value("Ab") - 0.8) < 1e-6
    assert p.value("size_margin") == float(p.size_margin)                                               # plain Python fields


def test_parameter_block_matches_oracle_constants():
    from larndsim_b200 import sim
    pp, op = cm.product_params(), cm.oracle_params()
    pod = sim.make_pod(pp)
    f = np.float32
    assert f(pod.vdrift) == f(oc.get_vdrift(op)) and pod.n_ticks == 2001 and pod.hold_interval == lo.hold_interval(op)
    assert f(pod.bin_width) == f(op.pixel_pitch / 10) and f(pod.efield_rho) == f(op.eField * op.lArDensity)
    sym, w = 2, op.pixel_pitch / 10
    assert np.array_equal(np.array(list(pod.tran_bin_edges), dtype=f), oc.linspace_jnp(f(-sym * w), f((sym + 1) * w), 6))
    assert np.array_equal(np.array(list(pod.long_diff_template)[:100], dtype=f), np.asarray(op.long_diff_template, dtype=f))
    assert np.allclose(np.array(pod.tpc_borders)[:2], np.asarray(op.tpc_borders, dtype=f))
    assert f(pod.ts_vdrift) == f(f(op.t_sampling) * f(oc.get_vdrift(op)))


def test_pad_size_matches_oracle_and_reference_rule():
    from larndsim_b200 import sim
    sim.size_history_dict.clear()
    hist = {}
    for n in (100, 103, 99, 250, 255, 262, 251, 1000, 1040, 1060):
        assert sim.pad_size(n, "t", 0.2) == lo.pad_size(n, "t", 0.2, hist)
    assert sim.pad_size(100, "fresh", 0.5) == 125 and sim.pad_size(110, "fresh", 0.5) == 125 and sim.pad_size(126, "fresh", 0.5) == 158
    assert sim.pad_size((10, 20), "nd", 0.1) == (11, 21)


def test_synthetic_generators():
    from larndsim_b200 import synthetic
    assert synthetic.FIELDS == cm.FIELDS
    a = synthetic.synthetic_response(7, 6, 300)
    assert np.array_equal(a, oc.synthetic_response(7, 6, 300))           # product and oracle generators are identical
    assert np.allclose(a[:5, :5].sum(-1) * 0.1, 1.0, atol=1e-5) and np.abs(a[5:].sum(-1)).max() < 1e-4
    tr, nev = synthetic.synthetic_tracks(20000, seed=3, precision=0.01)
    c = cm.FIELDS.index
    assert tr.shape == (20000, 26) and tr.dtype == np.float32 and nev == int(tr[:, c("eventID")].max()) + 1
    assert np.abs(tr[:, c("x")]).max() < 31.1 and np.abs(tr[:, c("y")]).max() < 62.1 and np.abs(tr[:, c("z")]).max() < 30.6
    assert np.allclose(tr[:, c("dE")], tr[:, c("dEdx")] * tr[:, c("dx")], rtol=1e-5)
    assert (np.diff(tr[:, c("eventID")]) >= 0).all()
    tr2, _ = synthetic.synthetic_tracks(20000, seed=3, precision=0.01)
    assert np.array_equal(tr, tr2)


def test_chop_tracks_conserves_energy_and_length():
    arr, _ = cm.fixture_batches(1, 0.01)[0]
    seg = np.load(os.path.join(cm.GOLD, "segments_input_1.npz"))["segments"]
    c = cm.FIELDS.index
    tot = sum(a[:, c("dx")].sum() for a, _ in cm.fixture_batches(1, 0.01))
    assert tot <= seg["dx"].sum() * (1 + 1e-4) and tot > 0.5 * seg["dx"].sum()     # over-long trajectories are dropped
    assert arr[:, c("dx")].max() <= 0.01 + 1e-6
    one = lo.structured_to_f32(lo.swap_xz_structured(seg[:1]))
    ch = lo.chop_tracks(one, cm.FIELDS, 0.01)
    assert abs(ch[:, c("dE")].sum() - one[0, c("dE")]) < 1e-4 * one[0, c("dE")]
    assert np.allclose(ch[-1, [c("x_end"), c("y_end"), c("z_end")]], one[0, [c("x_end"), c("y_end"), c("z_end")]])


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "larnd-sim-jax_b200", "larndsim_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from larndsim_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.LarndError):
        _lib.get_lib()


def test_cpu_tensors_are_rejected():
    import torch
    from larndsim_b200 import LarndError, sim
    with pytest.raises(LarndError):
        sim.simulate_wfs(cm.product_params(), torch.zeros(3, 5, 5, 10), torch.zeros(4, 26), cm.FIELDS)


def test_stage_operators_reject_cpu_tensors_and_bad_arguments(lib):
    """The stage-by-stage operators (quench, drift, simulate_signals, current_mc, accumulate_signals_parametrized) have no
    CPU path either, and their C entry points validate arguments before touching the device."""
    import torch
    from larndsim_b200 import LarndError, _lib, detsim, drifting, quenching, sim
    pp = cm.product_params(number_pix_neighbors=0)
    tr = torch.zeros(4, len(cm.FIELDS))
    for fn in (quenching.quench, drifting.drift):
        with pytest.raises(LarndError):
            fn(pp, tr, cm.FIELDS)
    with pytest.raises(LarndError):
        detsim.current_mc(pp, tr, torch.zeros(4, 2), cm.FIELDS)
    with pytest.raises(LarndError):
        detsim.accumulate_signals_parametrized(torch.zeros(3, 2001), torch.zeros(4, 51), torch.zeros(4, dtype=torch.int32),
                                               torch.zeros(4, dtype=torch.int32))
    with pytest.raises(LarndError):
        z = torch.zeros(4)
        sim.simulate_signals(pp, torch.zeros(3, dtype=torch.int32), z, z, torch.zeros(3, 5, 5, 10), z, z, z, z, z, z, z)
    pod, cols = sim.make_pod(pp), sim.make_columns(cm.FIELDS)
    assert lib.larnd_tracks_stage(None, 5, ctypes.byref(cols), None, ctypes.byref(pod), 7, None, None) == -1
    assert lib.larnd_current_mc(None, 5, None, None, ctypes.byref(pod), None, None, None) == -1
    assert lib.larnd_accumulate_parametrized(None, 3, 2001, None, 51, None, None, 4, None) == -1
    assert lib.larnd_signals_stream_forward(None, 0, None, None, None, None, None, 0, None, None, None, None, 0, ctypes.byref(pod), None,
                                            None, None, None) == -1
    assert b"larnd_signals_stream" in lib.larnd_last_error()
    # column tables name every column the stages touch
    oc = _lib.TrackColumns()
    assert {n for n, _ in oc._fields_} <= set(cm.FIELDS)


def test_numa_binding_helper_is_safe_without_topology(tmp_path):
    """bind_to_gpu_numa_node never raises: unknown topology (no GPU / no sysfs entry) -> None and the affinity is untouched."""
    import os
    from larndsim_b200 import parallel
    assert parallel._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert parallel._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    assert parallel.bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) is None
    assert os.sched_getaffinity(0) == before


def test_event_partition_is_balanced_and_complete():
    from larndsim_b200 import parallel
    rng = np.random.default_rng(0)
    ev = np.sort(rng.integers(0, 37, size=5000))
    ev[:11] = -1
    for world in (1, 2, 3, 8):
        parts = parallel.event_partition(ev, world)
        assert parts[0][0] == 0 and parts[-1][1] == 37 and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [int(((ev >= lo_) & (ev < hi)).sum()) for lo_, hi in parts]
        assert sum(sizes) == int((ev >= 0).sum()) and max(sizes) - min(sizes) < 2 * 5000 / 37 + 1


_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "larnd-sim-jax_b200")); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
from larndsim_b200 import parallel
import common as cm
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["MASTER_PORT"], rank=rank, world_size=world)
arr, _ = cm.fixture_batches(0, 0.01)[0]
local, nev, first = parallel.shard_tracks(arr, cm.FIELDS, rank, world)
ev = local[:, 0]
assert ev.min() == 0 and ev.max() == nev - 1
# every rank contributes its segment count, its dE sum and a fake 15-vector of gradients
red = torch.tensor([float(local.shape[0]), float(local[:, cm.FIELDS.index("dE")].sum())] + [float(rank + 1)] * 15, dtype=torch.float64)
parallel.allreduce_sum_(red)
assert int(red[0]) == arr.shape[0], (int(red[0]), arr.shape[0])
assert abs(float(red[1]) - float(arr[:, cm.FIELDS.index("dE")].sum())) < 1e-3
assert float(red[2]) == sum(range(1, world + 1))
g = parallel.allgather(torch.tensor([float(rank), float(nev)]))
assert g.shape == (world, 2) and int(g[:, 1].sum()) == int(arr[:, 0].max()) + 1
assert parallel.scan_points_for_rank(10, rank, world) == list(range(rank, 10, world))
dist.destroy_process_group()
print("ok", rank)
"""


def test_event_sharding_and_collectives_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = 29500 + os.getpid() % 2000
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out
        assert "ok" in out


_WORKER_AR = r"""
import os, sys
sys.path.insert(0, os.path.join(sys.argv[1], "larnd-sim-jax_b200"))
import torch, torch.distributed as dist
from larndsim_b200 import parallel
from larndsim_b200.losses import mse_adc
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["MASTER_PORT"], rank=rank, world_size=world)
torch.manual_seed(0)
n = 40
ev = torch.arange(n) // 10                      # 4 events, 2 per rank
Q0 = torch.rand(n, dtype=torch.float64) + 0.5
pos = torch.rand(n, 3, dtype=torch.float64) * 3
refQ = Q0 * 1.1; refpos = pos + 0.05
def loss_of(scale, sel, reduce):
    Q = Q0[sel] * scale
    l, _ = mse_adc(None, Q, pos[sel, 0], pos[sel, 1], pos[sel, 2], None, torch.ones_like(Q), ev[sel].double(),
                   refQ[sel], refpos[sel, 0], refpos[sel, 1], refpos[sel, 2], None, torch.ones_like(Q), ev[sel].double(), reduce=reduce)
    return l
s_all = torch.tensor(1.3, dtype=torch.float64, requires_grad=True)
full = loss_of(s_all, slice(None), None); full.backward()
s_loc = torch.tensor(1.3, dtype=torch.float64, requires_grad=True)
sel = (ev // 2) == rank
part = loss_of(s_loc, sel, parallel.allreduce_sum_differentiable); part.backward()
g = s_loc.grad.clone(); parallel.allreduce_sum_(g)
assert abs(float(part) - float(full)) < 1e-12, (float(part), float(full))
assert abs(float(g) - float(s_all.grad)) < 1e-10, (float(g), float(s_all.grad))
dist.destroy_process_group()
print("ok", rank)
"""


def test_sharded_loss_matches_single_process(tmp_path):
    """Two-phase reduction of mse_adc: loss and gradient from 2 event shards == single process."""
    script = tmp_path / "worker_ar.py"
    script.write_text(_WORKER_AR)
    port = 31500 + os.getpid() % 2000
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out


def test_host_side_key_split_matches_the_jax_restatement(lib):
    """larnd_rng_split runs on the host (no GPU needed): same keys as oracle/jax_random.split in both counter layouts."""
    from larndsim_b200 import jrandom
    from oracle import jax_random as jr
    for seed in (0, 7, 2 ** 40 + 3):
        for part in (True, False):
            for num in (1, 2, 5):
                got = jrandom.split(jrandom.key(seed), num, part)
                ref = [tuple(int(v) for v in k) for k in jr.split(jr.key(seed), num, part)]
                assert got == ref, (seed, part, num)


def test_event_id_validators_have_the_reference_error_behaviour():
    """detsim validators (detsim_jax.py:26-152): ValueError below -1 / non-contiguous local ids, OverflowError beyond the
    int64-safe packing limit, silent on empty or all-padding input; bin ids round-trip."""
    import torch
    from larndsim_b200 import detsim
    p = cm.product_params()
    detsim.validate_local_event_ids([])
    detsim.validate_local_event_ids([-1, -1])
    detsim.validate_local_event_ids([0, 1, -1, 2, 2])
    with pytest.raises(ValueError):
        detsim.validate_local_event_ids([0, 2])
    with pytest.raises(ValueError):
        detsim.validate_local_event_ids([-2, 0])
    detsim.validate_event_ids_for_packing(p, np.array([0, 5, -1]), kind="pixel", context="t")
    with pytest.raises(ValueError):
        detsim.validate_event_ids_for_packing(p, [-3], kind="pixel")
    with pytest.raises(ValueError):
        detsim.validate_event_ids_for_packing(p, [1], kind="voxel")
    lim = detsim.max_safe_event_id_for_pixel_packing(p)
    assert lim == (np.iinfo(np.int64).max - (140 * 280 * 2 - 1)) // (140 * 280 * 2)
    with pytest.raises(OverflowError):
        detsim.validate_event_ids_for_packing(p, [lim + 1], kind="pixel")
    assert detsim.max_possible_pixel_id(p, 3) == ((3 * 2 + 1) * 280 + 279) * 140 + 139
    detsim.validate_packed_ids_for_decoding(p, [0, 78400 * 7 + 5, -1])
    with pytest.raises(ValueError):
        detsim.validate_packed_ids_for_decoding(p, [-5])
    bx, by = torch.tensor([0, 1399, 17]), torch.tensor([2799, 0, 33])
    pl, ev = torch.tensor([0, 1, 1]), torch.tensor([0, 3, 9])
    bid = detsim.bin2id(p, bx, by, pl, ev)
    assert [t.tolist() for t in detsim.id2bin(p, bid)] == [bx.tolist(), by.tolist(), pl.tolist(), ev.tolist()]
    assert detsim.bin2id(p, torch.tensor([1400]), torch.tensor([0]), torch.tensor([0]), torch.tensor([0])).item() == -1


def test_bench_reference_arm_runs_on_cpu_and_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores) needs no GPU: one JSON line with the contract's keys."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, check=True).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "segments/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["steps"] == 1 and "workload" in line["config"]


def test_packing_validator_uses_the_int32_range_ids_are_packed_in():
    """pixel2id and the kernels pack ids in int32 (the reference never enables x64, so its astype(int64) is
    int32 too): an eventID whose packed id would wrap must be refused, not accepted against the int64 limit."""
    from larndsim_b200 import detsim
    p = cm.product_params()
    stride = 2 * p.n_pixels_x * p.n_pixels_y
    top_ok = (2 ** 31) // stride - 1
    detsim.validate_event_ids_for_packing(p, np.array([0, top_ok]), kind="pixel", context="test")
    with pytest.raises(OverflowError):
        detsim.validate_event_ids_for_packing(p, np.array([0, top_ok + 1]), kind="pixel", context="test")


def test_vdrift_follows_the_reference_arithmetic_for_static_and_fitted_efield():
    """Static eField: Python doubles rounded once; fitted eField (a traced float32 leaf in the reference): float32 op by
    op.  The derivative handed to the backward kernels is the closed form."""
    from larndsim_b200.consts import vdrift_and_derivative
    from oracle import consts as oc
    v_static, dv = vdrift_and_derivative(cm.product_params())
    v_leaf, dv2 = vdrift_and_derivative(cm.product_params(grad=("eField",)))
    assert v_static == oc.get_vdrift(cm.oracle_params())
    assert v_leaf == float(oc.get_vdrift(cm.oracle_params(), traced=True)) and v_leaf != v_static
    op = cm.oracle_params()
    h = 1e-6
    fd = (oc.get_vdrift(op.replace(eField=op.eField + h)) - oc.get_vdrift(op.replace(eField=op.eField - h))) / (2 * h)
    assert abs(dv - fd) < 1e-8 and dv == dv2


def test_parameter_block_is_cached_per_params_object_and_follows_in_place_updates():
    """sim.make_pod builds the C parameter block once per immutable Params object (it is requested several times per batch);
    the cache key carries the version counters of the tensor-valued fields, so an optimiser's in-place update is seen, and
    every call returns its own copy (callers patch fields of it)."""
    import torch
    import common as cm
    from larndsim_b200 import sim
    pp = cm.product_params(number_pix_neighbors=2, signal_length=150)
    a, b = sim.make_pod(pp, (32, 25, 25, 1950)), sim.make_pod(pp, (32, 25, 25, 1950))
    assert bytes(a) == bytes(b) and a is not b
    a.n_templates = 7
    assert sim.make_pod(pp, (32, 25, 25, 1950)).n_templates == b.n_templates != 7
    assert "_pod_cache" not in pp.replace(lifetime=100.0).__dict__          # replace() starts from the declared fields only
    assert abs(sim.make_pod(pp.replace(lifetime=100.0)).lifetime - 100.0) < 1e-6
    pg = cm.product_params(grad=("Ab", "lifetime"), number_pix_neighbors=2, signal_length=150)
    before = sim.make_pod(pg).lifetime
    with torch.no_grad():
        pg.lifetime.mul_(0.5)
    assert abs(sim.make_pod(pg).lifetime - 0.5 * before) < 1e-3 * before
    with pytest.raises(ValueError):
        sim.make_pod(pp, (1000, 25, 25, 1950))   # more bank rows than long_diff_template entries


def test_plain_float_params_carrier_builds_the_same_parameter_block():
    """consts.float_params_class: the value carrier of the fused fit step keeps the fitted-field list (a fitted eField means the
    float32 evaluation of the drift velocity, like the reference's traced scalar) but holds plain floats — its C parameter
    block must be byte-identical to the one built from the tensor-leaf Params object."""
    import torch
    import common as cm
    from larndsim_b200 import consts, sim
    names = ("Ab", "kb", "eField", "lifetime", "tran_diff", "long_diff")
    pg = cm.product_params(grad=names, number_pix_neighbors=2, signal_length=150)
    FC = consts.float_params_class(type(pg))
    assert FC is consts.float_params_class(type(pg)) and FC._grad_fields == type(pg)._grad_fields
    d = {k: getattr(pg, k) for k in consts._DEFAULTS}
    for k, v in d.items():
        if torch.is_tensor(v) and v.numel() == 1:
            d[k] = float(pg.value(k))
    pf = FC(**d)
    assert pf.grad_leaves() == [] and not any(torch.is_tensor(getattr(pf, n)) for n in names)
    assert bytes(sim.make_pod(pf)) == bytes(sim.make_pod(pg))
    # a fitted eField takes the float32 drift velocity in both; a static one the double-precision value rounded once
    ps = cm.product_params(number_pix_neighbors=2, signal_length=150)
    assert sim.make_pod(ps).vdrift == pytest.approx(sim.make_pod(pg).vdrift, rel=1e-6)
    pf2 = FC(**dict(d, lifetime=1234.5))
    assert abs(sim.make_pod(pf2).lifetime - 1234.5) < 1e-3
