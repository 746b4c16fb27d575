"""Shared set-up for the parity tests: module-0 parameters for the oracle and for the product, the prepared
fixture batches and the synthetic response bank (the real response_44.npy blob is absent from the reference)."""
import functools
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
for p in (ROOT, os.path.join(ROOT, "larnd-sim-jax_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import consts as oconsts  # noqa: E402
from oracle import larnd_oracle as lo  # noqa: E402

FIELDS = lo.FIELDS
GEOM = os.path.join(GOLD, "module0_geometry.json")


def oracle_params(**kw):
    base = dict(number_pix_neighbors=4, signal_length=100, electron_sampling_resolution=0.005,
                RESET_NOISE_CHARGE=0, UNCORRELATED_NOISE_CHARGE=0, time_window=100)
    base.update(kw)
    return oconsts.params_from_geometry_json(GEOM).replace(**base)


def product_params(grad=(), **kw):
    import larndsim_b200 as lb
    base = dict(number_pix_neighbors=4, signal_length=100, electron_sampling_resolution=0.005,
                RESET_NOISE_CHARGE=0, UNCORRELATED_NOISE_CHARGE=0, time_window=100)
    base.update(kw)
    cls = lb.build_params_class(list(grad))
    return lb.load_geometry_json(cls, GEOM).replace(**base)


@functools.lru_cache(maxsize=None)
def fixture_batches(ifile=0, precision=0.005, max_batch_len=50.0):
    """Chopped, event-localised batches of prepared_data/input_<ifile>.h5 (replay of TracksDataset)."""
    seg = np.load(os.path.join(GOLD, "segments_input_%d.npz" % ifile))["segments"]
    seg = lo.swap_xz_structured(seg)
    out = []
    for rows, gids in lo.make_batches(seg, max_batch_len):
        out.append((lo.batch_array(seg, rows, gids, FIELDS, True, precision), gids))
    return out


@functools.lru_cache(maxsize=None)
def synthetic_bank(n_templates=32, nx=45, ny=45, nt=1950):
    p = oracle_params()
    resp = oconsts.synthetic_response(nx, ny, nt)
    return oconsts.build_response_template(resp, p, n_templates=n_templates)


@functools.lru_cache(maxsize=None)
def synthetic_bank_cum(n_templates=32, nx=45, ny=45, nt=1950):
    return lo.response_cumsum(synthetic_bank(n_templates, nx, ny, nt))


def small_batch(n=600, ifile=0, ibatch=1, pad=40, precision=0.005):
    """A slice of a fixture batch + a few padding rows (eventID -1) as the reference's pad_batch makes them."""
    arr, _ = fixture_batches(ifile, precision)[ibatch]
    n = min(n, arr.shape[0])
    sub = arr[:n].copy()
    return lo.pad_batch(sub, n + pad, FIELDS) if pad else sub
