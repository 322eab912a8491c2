#!/usr/bin/env python
"""Compile the reference's URDFs / motion clips into small .npz assets.

/root/reference does not exist on the GPU box, so the static arrays the kernels need are
produced HERE by the package's own model compiler and committed under
ppr_diffphys_b200/assets/.  Re-run after changing ppr_diffphys_b200/model.py:

    python tools/compile_assets.py [/root/reference]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from ppr_diffphys_b200.model import compile_robot, ASSET_DIR  # noqa: E402


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    os.makedirs(ASSET_DIR, exist_ok=True)
    for name in ("laikago", "human", "quad"):
        m = compile_robot(name, os.path.join(ref, "data", "urdf_templates"))
        m.save(os.path.join(ASSET_DIR, name + ".npz"))
        print("%-8s nb=%d nq=%d nqd=%d nc=%d mass=%s" % (name, m.nb, m.nq, m.nqd, m.nc, np.round(m.body_mass, 3)))
        print("   parents", m.joint_parent.tolist())
    mdir = os.path.join(ref, "data", "motion_sequences")
    for seq in sorted(os.listdir(mdir)):
        with open(os.path.join(mdir, seq, "amp-%s.txt" % seq)) as fh:
            d = json.load(fh)
        frames = np.asarray(d["Frames"], dtype=np.float64)
        np.savez_compressed(os.path.join(ASSET_DIR, "motion_%s.npz" % seq), frames=frames.astype(np.float32),
                            frame_duration=np.float64(d["FrameDuration"]))
        print("motion %-14s frames=%s dt=%g" % (seq, frames.shape, d["FrameDuration"]))


if __name__ == "__main__":
    main()
