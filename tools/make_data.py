#!/usr/bin/env python
"""Derive the small committed occupancy fixtures from the reference's bundled
CT volumes (run once in the build container, where /root/reference exists).

data/<name>_gt40.bits.xz = xz(packbits(sample > 40)) over the 256^3 volume in
file order, i.e. byte (i*256+j)*256+k of <name>.dat is the sample at
(x=i, y=j, z=k) (reference: src/voxtrees/datareader.c:51-75) and the predicate
is the demo's `sample > threshold` with Threshold = 40
(examples/scene1.cfg:5, examples/scene2.cfg:4, src/demo/main.c:169).
The GPU box has no /root/reference, so the benchmark and the GPU tests read
these fixtures instead of the 16 MiB raw files.
"""
import hashlib
import lzma
import os
import sys

import numpy as np

REF = os.environ.get("VOX_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data")

for name in ("skull", "bonsai"):
    raw = np.fromfile(os.path.join(REF, "data", name + ".dat"), dtype=np.uint8)
    assert raw.size == 256 ** 3
    occ = raw > 40
    blob = lzma.compress(np.packbits(occ).tobytes(), preset=9 | lzma.PRESET_EXTREME)
    path = os.path.join(OUT, f"{name}_gt40.bits.xz")
    with open(path, "wb") as f:
        f.write(blob)
    print(name, int(occ.sum()), "voxels,", len(blob), "bytes, sha256(raw)=",
          hashlib.sha256(raw.tobytes()).hexdigest()[:16], file=sys.stderr)
