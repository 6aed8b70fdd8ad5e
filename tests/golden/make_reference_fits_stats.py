#!/usr/bin/env python
"""Generates tests/golden/reference_fits_stats.json from the reference's own FITS fixtures
(/root/reference/test/{sedov_rho_reference,sedov_T_reference,image}.fits) with the product's FITS reader.
Run in the dev container only (the GPU box has no /root/reference); the JSON is committed."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

s2g = ge.load_package()
from sphtogrid_b200 import io as s2gio

out = {}
for name in ("sedov_rho_reference.fits", "sedov_T_reference.fits", "image.fits"):
    path = os.path.join("/root/reference/test", name)
    hdus = s2gio.read_fits_hdus(path)
    hdr, img = hdus[0]
    out[name] = {"n_hdus": len(hdus), "header": {k: (v if not isinstance(v, float) else float(v)) for k, v in hdr.items()},
                 "shape": list(img.shape), "min": float(img.min()), "max": float(img.max()), "sum": float(img.sum()),
                 "corner": [float(img[0, 0]), float(img[0, -1]), float(img[-1, 0]), float(img[-1, -1])],
                 "center": float(img[img.shape[0] // 2, img.shape[1] // 2]),
                 "sha256_be_f64": hashlib.sha256(np.asfortranarray(img).astype(">f8").tobytes(order="F")).hexdigest()}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "reference_fits_stats.json"), "w"), indent=1, sort_keys=True)
print(json.dumps({k: (v["min"], v["max"], v["sum"]) for k, v in out.items()}))
