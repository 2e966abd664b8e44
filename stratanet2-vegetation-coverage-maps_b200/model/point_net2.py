"""B200-native ``model.point_net2`` -- drop-in for /root/reference/model/point_net2.py.

Same public names (``PointNet2``, ``SAModule``, ``GlobalSAModule``, ``FPModule``, ``MLP``), same
constructor arguments, same sub-module tree and therefore the same 53 ``state_dict`` keys / shapes
(SURVEY.md Appendix B), same ``save_state`` / ``load_state`` checkpoint dictionary, same
``forward(cloud_data)`` contract.  The arithmetic is not torch: ``forward`` runs the fused sm_100a
kernels of libsn2_b200.so (sn2/pipeline.py).  The nn.Module containers below only own parameters.

Eval mode runs the fully fused kernels; training mode (autograd) materialises the edge messages like the
reference (BatchNorm batch statistics need them), runs every graph operator and its backward in
libsn2_b200.so and leaves Linear/BatchNorm to the model's own torch modules (sn2/pipeline.py).

Not supported (raises, never falls back): CPU execution (``args.cuda is None`` models can be built, saved
and loaded, but not run) and ragged plots.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from sn2 import ops as _ops
from sn2 import pipeline as _pipeline


class PointConv(nn.Module):
    """Parameter holder standing in for torch_geometric.nn.PointConv (reference :19): keeps the
    ``local_nn`` attribute so that state_dict keys read ``...conv.local_nn.<i>.<j>.*``."""

    def __init__(self, local_nn=None, global_nn=None, add_self_loops=False):
        super().__init__()
        if add_self_loops or global_nn is not None:
            raise NotImplementedError("sn2 PointConv: only add_self_loops=False, global_nn=None (the reference setting)")
        self.local_nn = local_nn

    def forward(self, x, pos, edge_index):
        """Operator-level form (reference :27): message MLP over the edges + max aggregation, differentiable."""
        return _ops.pointconv(self.local_nn, x, pos, edge_index)


def MLP(channels, batch_norm=True):
    """Reference :45-53 -- a Sequential of (Linear, ReLU, BatchNorm1d) blocks, BN after ReLU."""
    blocks = []
    for cin, cout in zip(channels[:-1], channels[1:]):
        layers = [nn.Linear(cin, cout), nn.ReLU()]
        if batch_norm:
            layers.append(nn.BatchNorm1d(cout))
        blocks.append(nn.Sequential(*layers))
    return nn.Sequential(*blocks)


class SAModule(nn.Module):
    """Set abstraction: fps -> radius (cap 2000) -> PointConv max (reference :14-29)."""
    max_num_neighbors = 2000  # hard-coded at reference :24 (set the attribute on an instance to change the cap)

    def __init__(self, ratio, r, nn):
        super().__init__()
        self.ratio = ratio
        self.r = r
        self.conv = PointConv(nn, add_self_loops=False)

    def forward(self, x, pos, batch):
        """Operator-level form of reference :21-29 on (x, pos, batch) long-form tensors (dense batches).
        ``PointNet2.forward`` does not go through here: it runs the fused kernels (sn2/pipeline.py)."""
        idx = _ops.fps(pos, batch, ratio=self.ratio)
        row, col = _ops.radius(pos, pos[idx], self.r, batch, batch[idx], max_num_neighbors=self.max_num_neighbors)
        x = self.conv(x, (pos, pos[idx]), torch.stack([col, row], dim=0))
        return x, pos[idx], batch[idx]


class GlobalSAModule(nn.Module):
    """MLP on [x, pos] then per-plot max (reference :32-42)."""

    def __init__(self, nn):
        super().__init__()
        self.nn = nn

    def forward(self, x, pos, batch):
        """Operator-level form of reference :37-42."""
        x = _ops.global_max_pool(self.nn(torch.cat([x, pos], dim=1)), batch)
        return x, pos.new_zeros((x.size(0), 3)), torch.arange(x.size(0), device=batch.device)


class FPModule(nn.Module):
    """knn_interpolate -> concat skip -> MLP (reference :56-67)."""

    def __init__(self, k, nn):
        super().__init__()
        self.k = k
        self.nn = nn

    def forward(self, x, pos, batch, x_skip, pos_skip, batch_skip):
        """Operator-level form of reference :62-67."""
        x = _ops.knn_interpolate(x, pos, pos_skip, batch, batch_skip, k=self.k)
        if x_skip is not None:
            x = torch.cat([x, x_skip], dim=1)
        return self.nn(x), pos_skip, batch_skip


class PointNet2(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.cuda_device = args.cuda
        self.subsample_size = args.subsample_size
        self.n_class = args.n_class
        self.drop = args.drop
        self.n_input_feats = args.n_input_feats - 2  # x, y are not network features (reference :77)
        self.set_patience_attributes(args)
        self.log_embeddings = args.log_embeddings
        if self.n_input_feats != 8 or self.n_class != 4:
            raise NotImplementedError("sn2 PointNet2: kernels are built for n_input_feats=10, n_class=4 (reference defaults)")
        c_sa1 = [self.n_input_feats + 3, 16, 16]
        c_sa2 = [c_sa1[-1] + 3, 32]
        c_sa3 = [c_sa2[-1] + 3, 64]
        self.sa1_module = SAModule(args.ratio1, args.r1, MLP(c_sa1))
        self.sa2_module = SAModule(args.ratio2, args.r2, MLP(c_sa2))
        self.sa3_module = GlobalSAModule(MLP(c_sa3))
        c_fp3 = [c_sa3[-1] + c_sa2[-1], 64]
        c_fp2 = [c_fp3[-1] + c_sa1[-1], 34]
        c_fp1 = [c_fp2[-1] + self.n_input_feats, 34]
        self.fp3_module = FPModule(1, MLP(c_fp3))
        self.fp2_module = FPModule(3, MLP(c_fp2))
        self.fp1_module = FPModule(3, MLP(c_fp1))
        self.lin1 = nn.Linear(c_fp1[-1], 16)
        self.lin2 = nn.Linear(16, self.n_class + 1)
        # prior on [4 class logits, density logit] (reference :97-99)
        self.lin2.bias = nn.Parameter(torch.tensor([0.733, 0.266, 0.235, 0.358, 0.500]))
        self.softmax = nn.Softmax(dim=1)
        self.sigmoid = nn.Sigmoid()
        if self.cuda_device is not None:
            self.cuda(self.cuda_device)

    # ---------------------------------------------------------------------------------------
    def forward(self, cloud_data, trace=None, timer=None):
        """cloud_data: {"xyz": (B,3,N), "cloud": (B,10,N)} fp32 -> (coverages_pointwise, proba_pointwise),
        both (B*N,4) on the device, plot-major (reference :106-153)."""
        if self.cuda_device is None:
            raise RuntimeError("sn2 PointNet2.forward needs a CUDA device (args.cuda); this build has no CPU path")
        device = torch.device("cuda", self.cuda_device)
        with torch.cuda.device(device):
            if self.training:
                # also under torch.no_grad() (the reference permits it): batch statistics, running-stat updates,
                # no autograd graph
                cov, proba, g, cloud_dev = _pipeline.forward_train(
                    self, cloud_data["xyz"], cloud_data["cloud"], device, self.sa1_module.max_num_neighbors, trace, timer,
                    structure=cloud_data.get("sn2_structure"))
            else:
                cov, proba, g, cloud_dev = _pipeline.forward_eval(
                    self, cloud_data["xyz"], cloud_data["cloud"], device, self.sa1_module.max_num_neighbors, trace, timer)
        if self.log_embeddings:
            self.last_G_tensor = g
        # device copy of the normalised cloud, reused by model.project_to_2d to skip a second H2D
        self.last_cloud_device = cloud_dev
        return cov, proba

    def train(self, mode: bool = True):
        _pipeline.invalidate_packed(self)  # eval weights are BN-folded snapshots: never reuse one across a mode switch
        return super().train(mode)

    # -- layout helpers (reference :155-163) ---------------------------------------------------
    @staticmethod
    def get_long_form(data):
        """(B,f,N) -> (B*N,f)."""
        return data.permute(0, 2, 1).reshape(-1, data.shape[1])

    def get_batch_format(self, data):
        """(B*N,f) -> (B,f,N)."""
        return data.reshape(-1, self.subsample_size, data.shape[1]).transpose(1, 2)

    # -- early stopping bookkeeping (reference :165-184) ---------------------------------------
    def set_patience_attributes(self, args):
        self.stopped_early = False
        self.best_metric_value = 10 ** 6
        self.best_metric_epoch = 1
        self.patience_in_epochs = args.patience_in_epochs

    def stop_early(self, val_metric, epoch, args):
        if val_metric < self.best_metric_value:
            self.best_metric_value = val_metric
            self.best_metric_epoch = epoch
            self.save_state(args)
            return False
        if epoch < args.epoch_to_start_early_stop:
            return False
        if epoch >= self.best_metric_epoch + self.patience_in_epochs:
            self.stopped_early = True
            return True
        return False

    # -- checkpoints (reference :186-220) ---------------------------------------------------------
    @staticmethod
    def _checkpoint_path(args):
        tag = f"fold_n={args.current_fold_id}" if args.current_fold_id > 0 else "full"
        return os.path.join(args.stats_path, f"PCC_model_{tag}.pt")

    def save_state(self, args):
        torch.save(
            {
                "best_metric_epoch": self.best_metric_epoch,
                "state_dict": self.state_dict(),
                "best_metric_value": self.best_metric_value,
            },
            self._checkpoint_path(args),
        )

    def load_state(self, save_path):
        map_location = None if self.cuda_device is not None else torch.device("cpu")
        checkpoint = torch.load(save_path, map_location=map_location)
        self.load_state_dict(checkpoint["state_dict"])
        self.best_metric_epoch = checkpoint["best_metric_epoch"]
        self.best_metric_value = checkpoint["best_metric_value"]
        return self

    def load_best_state(self, args):
        return self.load_state(self._checkpoint_path(args))
