"""Deterministic synthetic knowledge-graph pairs shaped like DBP15K / DBP100K.

The reference reads DBP15K from disk (utils/data_utils.py:375-455), which is not
shipped; BASELINE.json's configs are synthetic pairs of the same shape
(SURVEY.md §8d).  This generator is host-side NumPy only and feeds the device
CSR builder and the benchmark; it takes no part in any timed region.
"""
from __future__ import annotations

import numpy as np

# name -> (E1, E2, triples1, triples2, rels1, rels2, links)
SHAPES = {
    "tiny": (300, 310, 1200, 1500, 17, 13, 200),
    "dbp15k": (19388, 19572, 70414, 95142, 1701, 1323, 15000),
    "dbp100k": (100000, 100000, 500000, 500000, 400, 300, 100000),
}


def _zipf_sample(rng, n_items, n_draws, exponent):
    """Draw ``n_draws`` ids with P(rank r) ∝ r^-exponent, ranks randomly assigned."""
    weights = np.arange(1, n_items + 1, dtype=np.float64) ** (-exponent)
    cdf = np.cumsum(weights)
    cdf /= cdf[-1]
    ranks = np.searchsorted(cdf, rng.random(n_draws), side="right")
    np.minimum(ranks, n_items - 1, out=ranks)
    return rng.permutation(n_items)[ranks]


def make_triples(rng, n_ent, n_triples, n_rel, offset=0, zipf=0.8):
    """[n_triples, 3] int64 (head, relation, tail); heads Zipf, tails uniform,
    every relation id used at least once (the reference's loader indexes
    head[rel] for every rel, utils/data_utils.py:403-404)."""
    heads = _zipf_sample(rng, n_ent, n_triples, zipf) + offset
    tails = rng.integers(0, n_ent, n_triples) + offset
    rels = rng.integers(0, n_rel, n_triples)
    rels[:min(n_rel, n_triples)] = np.arange(min(n_rel, n_triples))
    return np.stack([heads, rels, tails], 1).astype(np.int64)


def make_kg_pair(shape="dbp15k", dim=300, seed=0, noise=0.5, train_tenths=3, zipf=0.8,
                 features=True):
    """Synthetic pair.  KG1 ids are [0, E1), KG2 ids [E1, E1+E2).

    Returns dict: e1, e2, n, triples [T,3], links [L,2] (global ids), train, test
    (30/70 split by ``len//10*3`` like utils/data_utils.py:391-392), x [n, dim]
    fp32 row-L2-normalised with linked pairs correlated (x2 = x1 + noise·N(0,1)).
    """
    if isinstance(shape, str):
        e1, e2, t1, t2, r1, r2, n_links = SHAPES[shape]
    else:
        e1, e2, t1, t2, r1, r2, n_links = shape
    rng = np.random.default_rng(seed)
    kg1 = make_triples(rng, e1, t1, r1, 0, zipf)
    kg2 = make_triples(rng, e2, t2, r2, e1, zipf)
    kg2[:, 1] += r1
    n_links = min(n_links, e1, e2)
    left = rng.permutation(e1)[:n_links]
    right = rng.permutation(e2)[:n_links] + e1
    links = np.stack([left, right], 1).astype(np.int64)
    links = links[rng.permutation(n_links)]
    cut = n_links // 10 * train_tenths
    out = {"e1": e1, "e2": e2, "n": e1 + e2, "triples": np.concatenate([kg1, kg2]),
           "links": links, "train": links[:cut], "test": links[cut:], "n_rel": r1 + r2}
    if features:
        x = rng.standard_normal((e1 + e2, dim), dtype=np.float32)
        x[links[:, 1]] = x[links[:, 0]] + noise * rng.standard_normal((n_links, dim), dtype=np.float32)
        x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
        out["x"] = x.astype(np.float32)
    return out


def make_powerlaw_graph(n_nodes, avg_degree, seed=0, zipf=1.0):
    """SpMM-sweep graph (BASELINE.json config 4): ~n·avg_degree/2 triples with
    Zipf(1.0) heads, to be run through the same adjacency normalisation."""
    rng = np.random.default_rng(seed)
    n_triples = int(n_nodes * avg_degree / 2)
    heads = _zipf_sample(rng, n_nodes, n_triples, zipf)
    tails = rng.integers(0, n_nodes, n_triples)
    return heads.astype(np.int64), tails.astype(np.int64)


def write_dbp15k_dir(kg, root, lang="zh_en"):
    """Write a synthetic pair in the DBP15K on-disk layout the reference reads
    (utils/data_utils.py:375-384: ent_ids_{1,2}, rel_ids_{1,2}, triples_{1,2}, ref_ent_ids, ref_r_ids,
    <lang[:2]>_vectorList.json).  Returns the directory."""
    import json
    import os
    d = os.path.join(root, lang)
    os.makedirs(d, exist_ok=True)
    e1 = kg["e1"]
    tri = kg["triples"]
    kg1, kg2 = tri[tri[:, 0] < e1], tri[tri[:, 0] >= e1]
    rels1, rels2 = sorted(set(kg1[:, 1].tolist())), sorted(set(kg2[:, 1].tolist()))

    def dump(name, rows):
        with open(os.path.join(d, name), "w", encoding="utf-8") as f:
            for r in rows:
                f.write("\t".join(str(c) for c in r) + "\n")
    dump("ent_ids_1", [(i, "e%d" % i) for i in range(e1)])
    dump("ent_ids_2", [(i, "e%d" % i) for i in range(e1, kg["n"])])
    dump("rel_ids_1", [(r, "r%d" % r) for r in rels1])
    dump("rel_ids_2", [(r, "r%d" % r) for r in rels2])
    dump("triples_1", kg1.tolist())
    dump("triples_2", kg2.tolist())
    dump("ref_ent_ids", kg["links"].tolist())
    dump("ref_r_ids", [(rels1[0], rels2[0])])
    with open(os.path.join(d, lang[0:2] + "_vectorList.json"), "w", encoding="utf-8") as f:
        json.dump([[float(v) for v in row] for row in kg["x"]], f)
    return d
