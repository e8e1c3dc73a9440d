#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for lib in default $(ls variants/*.so); do
 for dbg in 0; do for NS in 8; do export NSETS=$NS;
  if [ "$lib" = default ]; then unset STITCHB200_LIB; else export STITCHB200_LIB=$PWD/$lib; fi
  SB_FS2_DEBUG=$dbg python - "$lib" $dbg <<'PY'
import sys, torch
import stitchingvideo_b200 as sv
from stitchingvideo_b200 import capi, rigs
Ks, Rs, spec = rigs.cameras("c2"); n = spec["n_used"]
comp = sv.Compositor((spec["W"], spec["H"]), Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=spec["blender"])
sets = [[capi.DeviceImage.from_torch(torch.from_numpy(rigs.frame("c2", s, i, smooth=0)).cuda()) for i in range(n)] for s in range(8)]
comp.set_depth(4)
import os
NSETS = int(os.environ.get("NSETS", "8"))
def run(k):
    slots = []
    for i in range(k):
        if i >= 4: comp.wait(slots[i - 4])
        slots.append(comp.enqueue(sets[i % NSETS], None))
    for s in slots[-4:]: comp.wait(s)
run(40)
comp.mark(0); run(400); comp.mark(1)
print("%-10s debug %s nsets %d: %.1f us per frame" % (sys.argv[1], sys.argv[2], NSETS, comp.marked_ms() * 1e3 / 400))
PY
 done; done
done
