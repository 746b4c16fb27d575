"""Synthetic inputs of the prepared_data shape and a synthetic response LUT.

The reference ships 22 tiny fixture files and its response LUT blobs are missing from the checkout, so the
benchmark workload is generated: events of a few straight tracks inside the module-0 TPCs, chopped into
``electron_sampling_resolution``-long segments with the arithmetic of the reference's chop_tracks
(optimize/dataio.py:63-106), in the file's column order (26 float32 columns, x/z already swapped so that z is
the drift axis).  Host-side numpy; nothing here is on the accelerated path.
"""
import math

import numpy as np

FIELDS = ("eventID", "z_end", "trackID", "tran_diff", "z_start", "x_end", "y_end", "n_electrons", "pdgId",
          "x_start", "y_start", "t_start", "t0_start", "t0_end", "t0", "dx", "long_diff", "pixel_plane", "t_end",
          "dEdx", "dE", "t", "y", "x", "z", "n_photons")


def synthetic_tracks(n_segments, seed=1234, precision=0.01, first_event=0):
    """(tracks (N,26) float32, n_events).  Generates whole events until at least ``n_segments`` rows exist and
    truncates at an event boundary when possible (the last event may be cut to hit N exactly)."""
    rng = np.random.default_rng(seed)
    c = {n: i for i, n in enumerate(FIELDS)}
    chunks = []
    total = 0
    ev = first_event
    while total < n_segments:
        ntracks = int(rng.integers(1, 5))
        for trk in range(ntracks):
            sign = 1.0 if rng.random() < 0.5 else -1.0
            start = np.array([rng.uniform(-30, 30), rng.uniform(-61, 61), sign * rng.uniform(0.5, 30.0)])
            while True:
                d = rng.normal(size=3)
                d /= np.linalg.norm(d)
                if abs(d[2]) < 0.966:
                    break
            length = rng.uniform(10, 60)
            # clip to the TPC volume (x,y borders and the drift volume of the track's own TPC)
            lims = []
            for k, (lo, hi) in enumerate(((-30.9, 30.9), (-61.9, 61.9), (0.2, 30.5) if sign > 0 else (-30.5, -0.2))):
                if d[k] > 0:
                    lims.append((hi - start[k]) / d[k])
                elif d[k] < 0:
                    lims.append((lo - start[k]) / d[k])
            length = max(min([length] + lims), precision)
            nst = max(int(math.ceil(length / precision)), 1)
            steps = np.arange(nst, dtype=np.float64)
            dedx = float(np.exp(rng.uniform(math.log(1.5), math.log(25.0))))
            t0 = rng.uniform(1e-4, 3e-3)
            seg = np.zeros((nst, len(FIELDS)), dtype=np.float32)
            for k, ax in enumerate("xyz"):
                s0 = start[k] + steps * precision * d[k]
                s1 = start[k] + precision * (steps + 1) * d[k]
                s1[-1] = start[k] + length * d[k]
                seg[:, c[ax + "_start"]] = s0
                seg[:, c[ax + "_end"]] = s1
                seg[:, c[ax]] = 0.5 * (seg[:, c[ax + "_start"]] + seg[:, c[ax + "_end"]])
            dx = np.full(nst, precision)
            dx[-1] = length - precision * (nst - 1)
            seg[:, c["dx"]] = dx
            seg[:, c["dEdx"]] = dedx
            seg[:, c["dE"]] = dedx * dx
            seg[:, c["eventID"]] = ev - first_event
            seg[:, c["trackID"]] = trk
            seg[:, c["pdgId"]] = 13
            seg[:, c["t0"]] = t0
            seg[:, c["t0_start"]] = t0
            seg[:, c["t0_end"]] = t0
            chunks.append(seg)
            total += nst
        ev += 1
    tracks = np.concatenate(chunks, axis=0)[:n_segments]
    n_events = int(tracks[:, c["eventID"]].max()) + 1
    return np.ascontiguousarray(tracks, dtype=np.float32), n_events


def synthetic_raw_tracks(n_segments, seed=1234, precision=0.01, first_event=0):
    """(raw (M,26) float32, n_events): the same events as synthetic_tracks but ONE ROW PER TRACK (un-chopped, the shape of
    the prepared_data files); chopping them at ``precision`` (dataio.chop_tracks on the device, or the reference's
    chop_tracks) yields at least ``n_segments`` rows."""
    rng = np.random.default_rng(seed)
    c = {n: i for i, n in enumerate(FIELDS)}
    rows = []
    total = 0
    ev = first_event
    while total < n_segments:
        ntracks = int(rng.integers(1, 5))
        for trk in range(ntracks):
            sign = 1.0 if rng.random() < 0.5 else -1.0
            start = np.array([rng.uniform(-30, 30), rng.uniform(-61, 61), sign * rng.uniform(0.5, 30.0)])
            while True:
                d = rng.normal(size=3)
                d /= np.linalg.norm(d)
                if abs(d[2]) < 0.966:
                    break
            length = rng.uniform(10, 60)
            lims = []
            for k, (lo, hi) in enumerate(((-30.9, 30.9), (-61.9, 61.9), (0.2, 30.5) if sign > 0 else (-30.5, -0.2))):
                if d[k] > 0:
                    lims.append((hi - start[k]) / d[k])
                elif d[k] < 0:
                    lims.append((lo - start[k]) / d[k])
            length = max(min([length] + lims), precision)
            dedx = float(np.exp(rng.uniform(math.log(1.5), math.log(25.0))))
            t0 = rng.uniform(1e-4, 3e-3)
            row = np.zeros(len(FIELDS), dtype=np.float32)
            for k, ax in enumerate("xyz"):
                row[c[ax + "_start"]] = start[k]
                row[c[ax + "_end"]] = start[k] + length * d[k]
                row[c[ax]] = 0.5 * (row[c[ax + "_start"]] + row[c[ax + "_end"]])
            row[c["dx"]] = length
            row[c["dEdx"]] = dedx
            row[c["dE"]] = dedx * length
            row[c["eventID"]] = ev - first_event
            row[c["trackID"]] = trk
            row[c["pdgId"]] = 13
            row[c["t0"]] = row[c["t0_start"]] = row[c["t0_end"]] = t0
            rows.append(row)
            total += max(int(math.ceil(length / precision)), 1)
        ev += 1
    raw = np.stack(rows, axis=0)
    return np.ascontiguousarray(raw, dtype=np.float32), ev - first_event


def synthetic_response(nx=45, ny=45, nt=1950, seed=7, t_sampling=0.1):
    """Synthetic stand-in for the missing response_44.npy: collecting bins (i,j < 5) carry a unipolar pulse
    (tau ~ 12 ticks, peak ~70 ticks before the end of the axis, one electron = unit integral), the others a
    zero-net bipolar induction pulse decaying with max(i,j).  Same construction as the oracle's generator
    (tests assert they are identical)."""
    rng = np.random.default_rng(seed)
    t = np.arange(nt, dtype=np.float64)
    peak = nt - 70.0
    resp = np.zeros((nx, ny, nt), dtype=np.float64)
    jitter = rng.uniform(-1.0, 1.0, size=(nx, ny))
    for i in range(nx):
        for j in range(ny):
            r = max(i, j)
            d = math.hypot(i, j)
            pk = peak + 0.6 * jitter[i, j] - 0.15 * d
            if i < 5 and j < 5:
                tau = 12.0 + 0.8 * d
                rise = np.where(t <= pk, np.exp(np.minimum(t - pk, 0.0) / tau), np.exp(-np.maximum(t - pk, 0.0) / 1.5))
                rise /= rise.sum() * t_sampling
                resp[i, j] = rise
            else:
                amp = 0.08 * math.exp(-(r - 4) / 6.0) * (1.0 + 0.1 * jitter[i, j])
                w = 14.0 + 0.5 * r
                gp = np.exp(-0.5 * ((t - (pk - 1.2 * w)) / w) ** 2)
                gm = np.exp(-0.5 * ((t - (pk + 0.2 * w)) / (0.5 * w)) ** 2)
                gp /= gp.sum()
                gm /= gm.sum()
                resp[i, j] = amp * (gp - gm) / t_sampling * 0.1
    return resp.astype(np.float32)
