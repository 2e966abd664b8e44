"""ncu target: parcel grid + plot extraction + band finalisation on a small synthetic parcel (side metres, density pts/m^2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from sn2.parcel import ParcelCloud, extract_plots, finalize_mosaic, plot_centers_reference  # noqa: E402

side = float(sys.argv[1])
dev = torch.device("cuda", 0)
args, net = bench.make_model(10000, 0)
cloud = bench.synthetic_parcel(dev, side=side)
for _ in range(1 if os.environ.get("SN2_NCU_ONE") == "1" else 2):
    parcel = ParcelCloud(cloud, dev)
    centers = plot_centers_reference(parcel.x_min, parcel.x_max, parcel.y_min, parcel.y_max, args)
    ex = extract_plots(parcel, centers, args)
    m = torch.rand(4, 1058, 1058, dtype=torch.float64, device=dev)
    m[:, torch.rand(1058, 1058, device=dev) < 0.1] = float("nan")
    out, thr, _ = finalize_mosaic(m)
    torch.cuda.synchronize()
print("ok", centers.shape[0], int(ex["valid"].sum()), float(thr))
