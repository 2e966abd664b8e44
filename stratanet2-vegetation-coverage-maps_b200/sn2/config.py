"""The argparse fields the hot path reads, with the reference defaults.

Mirrors the subset of /root/reference/config.py:21-101 that ``PointNet2(args)`` and the two
projection functions consume: the product accepts the reference's own ``args``
Namespace unchanged; this helper only exists so tests and bench.py can build one without argparse.
"""
from argparse import Namespace

import numpy as np


def default_args(**overrides) -> Namespace:
    args = Namespace(
        cuda=None,                      # config.py:21
        n_class=4,                      # config.py:52
        n_input_feats=10,               # config.py:101 (len(FEATURE_NAMES))
        subsample_size=10000,           # config.py:67
        diam_meters=20,                 # config.py:68
        diam_pix=20,                    # config.py:69
        drop=0.0,                       # config.py:76
        ratio1=0.25,                    # config.py:77
        r1=float(np.sqrt(2.0)),         # config.py:78
        ratio2=0.25,                    # config.py:79
        r2=float(np.sqrt(8.0)),         # config.py:80
        patience_in_epochs=30,          # config.py:92
        epoch_to_start_early_stop=250,  # config.py:90
        log_embeddings=False,           # config.py:41
        m=0.10,                         # config.py:70  weight of the NLL term
        e=0.2 / 5,                      # config.py:71  weight of the entropy term
        z_max=24.24,                    # config.py:73
        znorm_radius_in_meters=1.5,     # config.py:72
        lr=1e-3, wd=0.001, step_size=1, lr_decay=0.985, batch_size=20,  # config.py:84-98
        current_fold_id=0,
        stats_path=".",
    )
    for k, v in overrides.items():
        setattr(args, k, v)
    return args
