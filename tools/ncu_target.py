"""Small ncu target: one eval forward + projections (after warm-up launches) at N points per plot.
usage: ncu ... python tools/ncu_target.py N B K tensor_core [train]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200")]
import torch  # noqa: E402

from bench import make_model  # noqa: E402
from sn2 import ops  # noqa: E402
from sn2.pipeline import forward_eval  # noqa: E402
from sn2.synth import synth_batch  # noqa: E402

N, B, K, tc = (int(a) for a in sys.argv[1:5])
dev = torch.device("cuda", 0)
args, net = make_model(N, 0)
net.sa1_module.max_num_neighbors = K
net.sn2_tensor_core = tc
data = {k: v.to(dev) for k, v in synth_batch(5 if K < 2000 else 2, B, N).items()}
with torch.no_grad():
    for _ in range(1 if os.environ.get("SN2_NCU_ONE") == "1" else 2):
        cov, proba, g, cloud_d = forward_eval(net, data["xyz"], data["cloud"], dev, K)
        ops.project_plotwise(cloud_d, cov, args.diam_pix)
        ops.project_rasters(cloud_d, cov, "point_major", args.diam_pix, args.diam_meters)
        torch.cuda.synchronize()
print("ok")
