"""Target program for ncu captures: builds the C2 engine (ResNet-50-FPN, B=64, beam 8, V=10000) with a short decode
(max_len 8) and runs one untimed generate so that every kernel of the path is launched a few times.
    ncu --set full ... python tests/ncu_driver.py [workload] [max_len]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fpn-mt-image-captioning_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from fpnmt.engine import Engine  # noqa: E402
from fpnmt.weights import init_weights  # noqa: E402

wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"])
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
w = init_weights(wl["backbone"], vocab=wl["vocab"], seed=0)
eng = Engine(w, backbone=wl["backbone"], batch=wl["batch"], beam=wl["beam"], vocab=wl["vocab"], max_len=T, use_graphs=False)
img = torch.rand(wl["batch"], 512, 512, 3, generator=torch.Generator().manual_seed(0)).cuda() * 2 - 1
ids, lens = eng.generate(img, early_stop=False)
torch.cuda.synchronize()
print("ok", int(ids.sum()), eng.launch_count)
