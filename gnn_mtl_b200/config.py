"""Defaults of the reference's config.py:3-35 as a plain namespace factory (the reference's argparse
wiring crashes on any CLI override — SURVEY.md §5 — so values are edited in code there too)."""
import types

DEFAULTS = {
    'unsup': True, 'lr': 0.001, 'dropout': 0.0, 'cuda': 0, 'epochs': 100, 'weight_decay': 0.0, 'seed': 10086,
    'task': 'ea', 'model': 'GCN', 'num_layers': 3, 'act': 'relu', 'dim': 300, 'n_heads': 4, 'alpha': 0.2,
    'dataset': 'zh_en', 'normalize_x': 0, 'normalize_adj': 1, 'patience': 10, 'log_freq': 1, 'eval_freq': 1,
    'lr_reduce_freq': 2000, 'gamma': 0.5, 'min_epochs': 100, 'use_feats': 1, 'bias': 1, 'neg_num': 125,
    'batch_size': 3000, 'save': 0, 'iters': 1, 'refine_epochs': 5, 'refine_size': 3000,
}


def make_args(**overrides):
    unknown = set(overrides) - set(DEFAULTS) - {"data_root"}
    if unknown:
        raise KeyError("unknown config keys: %s" % sorted(unknown))
    return types.SimpleNamespace(**{**DEFAULTS, "data_root": "data/dbp15k", **overrides})
