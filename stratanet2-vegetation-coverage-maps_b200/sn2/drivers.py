"""Batched evaluation and pseudo-labelling on top of the batched kernels.

The reference evaluates with ``DataLoader(batch_size=1)`` (learning/test.py:38-42): one forward, one projection
(``torch.unique`` on the CPU), three losses and six ``.item()`` synchronisations per plot; its pseudo-labelling pass
(predict.py:104-111) projects batch by batch but copies every batch to the host.  The functions here run B plots per
launch and synchronise once at the end; the numbers they return are the reference's -- the meters of ``evaluate`` are
averages of PER-PLOT losses, so every loss is reduced per plot first.
"""
from __future__ import annotations

import torch

from . import losses

MIN_POINTS_NB_FOR_PSEUDO_LABELLING = 2000  # inference/predict_utils.py:64


def per_plot_losses(pred_pl, gt, proba, pdf, N: int, m: float, e: float):
    """The quantities ``evaluate`` feeds its meters with (learning/test.py:61-76), for every plot of a batch:
    -> dict of (B,) tensors: total_loss, MAE_loss, log_loss, MAE_veg_b, MAE_veg_moy, MAE_veg_h."""
    B = pred_pl.shape[0]
    sel = lambda t: torch.cat([t[:, :1], t[:, 2:4]], dim=1)  # noqa: E731
    by_strata = ((sel(pred_pl) - sel(gt)).pow(2) + losses.EPS).pow(0.5)                        # (B,3): get_absolute_loss_by_strata per plot
    mae = by_strata.mean(1)
    p = proba.view(B, N, 4)
    f = pdf.view(B, N, 3)
    lik = (p[:, :, 0] + p[:, :, 1]).double() * f[:, :, 0] + p[:, :, 2].double() * f[:, :, 1] + p[:, :, 3].double() * f[:, :, 2]
    nll = -torch.log(lik).mean(1)                                                               # get_NLL_loss per plot (float64)
    q = p[:, :, 2:]
    ent = -(q * torch.log(q + losses.EPS) + (1 - q) * torch.log(1 - q + losses.EPS)).mean(dim=(1, 2))  # get_entropy_loss per plot
    return {"total_loss": mae + m * nll + e * ent, "MAE_loss": mae, "log_loss": nll,
            "MAE_veg_b": by_strata[:, 0], "MAE_veg_moy": by_strata[:, 1], "MAE_veg_h": by_strata[:, 2]}


@torch.no_grad()
def evaluate_batched(model, batches, args, kde_lut, want_predictions: bool = True):
    """``evaluate`` (learning/test.py:24-132) without its plotting / Comet side: model.eval(), every batch
    {"xyz", "cloud", "coverages"} through forward + project_to_plotwise_coverages + the three losses.
    -> (dict of meter averages as the reference returns them, list of per-plot (pred [4], gt [4]) when asked).
    One host synchronisation at the end."""
    from model.project_to_2d import project_to_plotwise_coverages

    model.eval()
    acc, preds, n_plots = None, [], 0
    for batch in batches:
        cov, proba = model({"xyz": batch["xyz"], "cloud": batch["cloud"]})
        cloud_d = model.last_cloud_device
        B, _, N = cloud_d.shape
        pred_pl = project_to_plotwise_coverages(cov, cloud_d, args)
        gt = batch["coverages"].to(pred_pl.device, non_blocking=True).to(pred_pl.dtype)
        pdf = kde_lut.pdf(cloud_d, args.z_max)
        per = per_plot_losses(pred_pl, gt, proba, pdf, N, args.m, args.e)
        sums = torch.stack([v.double().sum() for v in per.values()])
        acc = sums if acc is None else acc + sums
        n_plots += B
        if want_predictions:
            preds.append((pred_pl, gt))
    if acc is None:
        raise RuntimeError("evaluate_batched: empty dataset")
    means = (acc / n_plots).cpu().tolist()  # the one synchronisation
    out = dict(zip(("total_loss", "MAE_loss", "log_loss", "MAE_veg_b", "MAE_veg_moy", "MAE_veg_h"), means))
    out["step"] = getattr(args, "current_step_in_fold", 0)
    summaries = []
    if want_predictions:
        for pred_pl, gt in preds:
            summaries.extend(zip(pred_pl.cpu().numpy(), gt.cpu().numpy()))
    return out, summaries


@torch.no_grad()
def pseudo_label(model, dataset: dict, collate, args, batch_size: int = 64) -> dict:
    """predict.py:104-111 + filter_dataset (inference/predict_utils.py:62-71): plots with more than 2000 raw points get a
    ``"coverages"`` entry = their predicted plot-wise coverages [4].  dataset: {plot_id: cloud_data with
    "N_points_in_cloud"}; collate(list of cloud_data) -> {"xyz": (B,3,N), "cloud": (B,10,N)} (the DataLoader's job).
    All batches are enqueued first; the predictions come back in one copy."""
    from model.project_to_2d import project_to_plotwise_coverages

    model.eval()
    kept = {k: v for k, v in dataset.items() if v["N_points_in_cloud"] > MIN_POINTS_NB_FOR_PSEUDO_LABELLING}
    ids = list(kept)
    outs = []
    for b0 in range(0, len(ids), batch_size):
        batch = collate([kept[i] for i in ids[b0:b0 + batch_size]])
        cov, _ = model(batch)
        outs.append(project_to_plotwise_coverages(cov, model.last_cloud_device, args))
    if outs:
        pred = torch.cat(outs).cpu().numpy()
        for i, plot_id in enumerate(ids):
            kept[plot_id].update({"coverages": pred[i].squeeze()})
    return kept
