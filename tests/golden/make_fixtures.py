"""Generates the committed fixtures under tests/golden/ from the read-only reference checkout.

Run once in the build container (where /root/reference exists):
    python tests/golden/make_fixtures.py

Outputs (all small, committed):
  module0_geometry.json      derived geometry/constants from module0.yaml + multi_tile_layout-2.4.16_v4.yaml
  segments_input_{0..21}.npz the raw `segments` records of the 22 un-shifted prepared_data/input_*.h5 files (structured
                             array, unswapped) = BASELINE config 2
  golden_lut_{0..4}.npz      per-batch/per-event golden hits of output/jax_ref/output_{0..4}.h5
                             (keys "b<batch>/e<event>/<dataset>")
/root/reference does not exist on the GPU box; tests read only these fixtures.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import consts  # noqa: E402
from oracle.h5lite import H5Lite  # noqa: E402

REF = "/root/reference"


def main():
    p = consts.load_detector_properties(
        os.path.join(REF, "src/larndsim/detector_properties/module0.yaml"),
        os.path.join(REF, "src/larndsim/pixel_layouts/multi_tile_layout-2.4.16_v4.yaml"))
    consts.save_geometry_json(p, os.path.join(HERE, "module0_geometry.json"))
    # the package ships the same derived geometry for bench.py / smoke() (no YAML on the GPU box)
    consts.save_geometry_json(p, os.path.join(HERE, "..", "..", "larnd-sim-jax_b200", "larndsim_b200", "data", "module0_geometry.json"))
    for i in range(22):
        seg = H5Lite(os.path.join(REF, "prepared_data/input_%d.h5" % i)).read("/segments")
        np.savez_compressed(os.path.join(HERE, "segments_input_%d.npz" % i), segments=seg)
        if i >= 5:   # goldens exist for files 0-4 only (output/jax_ref)
            print(i, seg.shape)
            continue
        g = H5Lite(os.path.join(REF, "output/jax_ref/output_%d.h5" % i))
        out = {}
        for b in g.keys("/"):
            for e in g.keys("/" + b):
                for ds in g.keys("/%s/%s" % (b, e)):
                    out["%s/%s/%s" % (b.replace("batch_", "b"), e.replace("event_", "e"), ds)] = g.read("/%s/%s/%s" % (b, e, ds))
        np.savez_compressed(os.path.join(HERE, "golden_lut_%d.npz" % i), **out)
        print(i, seg.shape, len(out))


if __name__ == "__main__":
    main()
