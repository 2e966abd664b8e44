"""Generate the golden vectors of tests/golden/*.npz.

Runs the REFERENCE's own files (/root/reference/model/point_net2.py and model/project_to_2d.py,
imported verbatim through oracle/ref_loader.py) on top of the restated third-party ops, on small
seeded synthetic plots, and stores inputs, weights and outputs.  /root/reference only exists in the
build container, so the vectors are committed; this script is how they were made:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200"))

from oracle import thirdparty_ops as tp  # noqa: E402
from oracle.ref_loader import load_reference, reference_root  # noqa: E402
from sn2.config import default_args  # noqa: E402
from sn2.synth import randomize_bn_, synth_batch  # noqa: E402

CASES = {
    "case_plain_b2_n1024": dict(config=7, B=2, N=1024, variant="plain"),
    "case_dup_b1_n2048": dict(config=8, B=1, N=2048, variant="dup"),
    "case_cm_b2_n1536": dict(config=9, B=2, N=1536, variant="cm"),
}


def main():
    assert reference_root() == "/root/reference", "golden vectors must come from the real reference tree"
    pn2, p2d = load_reference()
    for name, c in CASES.items():
        args = default_args(subsample_size=c["N"])
        torch.manual_seed(0)
        net = pn2.PointNet2(args)
        randomize_bn_(net)
        net.eval()
        data = synth_batch(c["config"], c["B"], c["N"], c["variant"])
        with torch.no_grad():
            cov, proba = net(data)
            plotwise = p2d.project_to_plotwise_coverages(cov, data["cloud"], args)
            cov_b = net.get_batch_format(cov)
            rasters = np.stack([p2d.project_to_2d_rasters(data["cloud"][b], cov_b[b], args) for b in range(c["B"])])
        pos = data["xyz"].permute(0, 2, 1).reshape(-1, 3)
        batch = torch.arange(c["B"]).repeat_interleave(c["N"])
        idx1 = tp.fps(pos, batch, ratio=args.ratio1)
        e1 = tp.radius(pos, pos[idx1], args.r1, batch, batch[idx1], max_num_neighbors=2000)
        out = {
            "xyz": data["xyz"].numpy(), "cloud": data["cloud"].numpy(),
            "cov": cov.numpy(), "proba": proba.numpy(), "plotwise": plotwise.numpy(), "rasters": rasters,
            "idx1": idx1.numpy(), "edges1": e1.numpy().astype(np.int32),
        }
        for k, v in net.state_dict().items():
            out["sd:" + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if not k.startswith("sd:")})


TRAIN_CASES = {
    "train_plain_b2_n2048": dict(config=3, B=2, N=2048, variant="plain"),
}


def train_loss(proba, pw, gt, pdf):
    """The loss the training parity tests use (shape of learning/loss_functions.py: MAE on three coverages, a
    likelihood term on the class probabilities, an entropy term), written with plain torch ops."""
    mae = torch.sqrt((pw[:, [0, 2, 3]] - gt[:, [0, 2, 3]]) ** 2 + 1e-4).mean()
    nll = -torch.log((proba[:, [0, 2, 3]].double() * pdf).sum(1) + 1e-6).mean().float()
    p = proba[:, 2:]
    ent = -(p * torch.log(p + 1e-6)).sum(1).mean()
    return mae + 0.10 * nll + 0.04 * ent


def main_train():
    """Train-mode vectors: the reference's own PointNet2 under model.train() (BatchNorm batch statistics, dropout
    switched off so that the step is deterministic), plot-wise projection, loss, backward.  Stored: inputs, initial
    state_dict, loss, all 32 parameter gradients, the state_dict after the forward (running statistics)."""
    assert reference_root() == "/root/reference", "golden vectors must come from the real reference tree"
    pn2, p2d = load_reference()
    for name, c in TRAIN_CASES.items():
        args = default_args(subsample_size=c["N"])
        args.drop = 0.0
        torch.manual_seed(0)
        net = pn2.PointNet2(args)
        randomize_bn_(net)
        net.drop = 0.0
        net.train()
        out = {}
        for k, v in net.state_dict().items():
            out["sd:" + k] = v.numpy().copy()
        data = synth_batch(c["config"], c["B"], c["N"], c["variant"])
        g = torch.Generator().manual_seed(9)
        gt = torch.rand(c["B"], 4, generator=g)
        z = data["xyz"][:, 2, :].reshape(-1, 1).double()
        pdf = torch.cat([torch.exp(-z), 0.5 * torch.exp(-0.5 * (z - 1.0) ** 2), 0.1 + 0.05 * z], dim=1)
        cov, proba = net(data)
        pw = p2d.project_to_plotwise_coverages(cov, data["cloud"], args)
        loss = train_loss(proba, pw, gt, pdf)
        loss.backward()
        out.update({"xyz": data["xyz"].numpy(), "cloud": data["cloud"].numpy(), "gt": gt.numpy(), "pdf": pdf.numpy(),
                    "loss": loss.detach().numpy(), "cov": cov.detach().numpy(), "plotwise": pw.detach().numpy()})
        for k, p in net.named_parameters():
            out["grad:" + k] = p.grad.numpy().copy()
        for k, v in net.state_dict().items():
            out["after:" + k] = v.numpy().copy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "loss", float(loss), "grads", sum(1 for k in out if k.startswith("grad:")))


if __name__ == "__main__":
    if "--train" in sys.argv:
        main_train()
    else:
        main()
        main_train()
