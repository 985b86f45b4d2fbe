#!/usr/bin/env python
"""Generates the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref built by oracle/build_ref.py):

    python tests/golden/make_golden.py

Outputs (committed):
  ref_labels.npz      label arrays produced by the reference's own ``Clustering.fit`` for the cases
                      listed in ``cases.py`` (inputs are regenerated from seeds, never stored)
  ref_labels_recipes.npz  the same for ``cases.RECIPE_CASES`` ("distances" and "neighbourhoods" recipes)
  records_sample.json a distilled sample of the reference's vendored parameter-scan records
                      (docsrc/benchmark/records/*.json): n_clusters / ratio_largest / ratio_noise
                      for (data set, r, c) cells whose data can be regenerated from code
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref  # noqa: E402
import cases  # noqa: E402

REF_ROOT = "/root/reference"


def main():
    assert ref.available(), "build oracle/_ref first (python oracle/build_ref.py)"
    out = {}
    for name, case in cases.CASES.items():
        x = case["data"]()
        kw = dict(case["fit"])
        metric = kw.pop("metric", "euclidean_r")
        extra = {} if metric == "euclidean_r" else {"fitter__ngetter__dgetter__metric": metric}
        labels = ref.fit_points(x, kw.pop("r"), kw.pop("c"), **kw, **extra)
        out[name] = labels.astype(np.int32)
        print(f"{name:28s} n={x.shape[0]:6d} clusters={int(labels.max()):4d} noise={(labels == 0).mean():.3f}")
    np.savez_compressed(os.path.join(HERE, "ref_labels.npz"), **out)

    # the "distances" and "neighbourhoods" recipes of the reference (recipes.py:329-348)
    out = {}
    mod = ref.module()
    for name, case in cases.RECIPE_CASES.items():
        data = case["data"]()
        kw = dict(case["fit"])
        cl = mod.Clustering(data, recipe=case["recipe"])
        cl.fit(kw.pop("r"), kw.pop("c"), v=False, **kw)
        labels = np.array(cl.labels, dtype=np.int64)
        out[name] = labels.astype(np.int32)
        print(f"{name:28s} n={labels.shape[0]:6d} clusters={int(labels.max()):4d} noise={(labels == 0).mean():.3f}")
    np.savez_compressed(os.path.join(HERE, "ref_labels_recipes.npz"), **out)

    # vendored scan records: keep a deterministic sample per reproducible set
    sample = {}
    rng = np.random.default_rng(0)
    for name in ("no_structure", "blobs", "varied", "aniso", "circles"):
        with open(os.path.join(REF_ROOT, "docsrc/benchmark/records", name + ".json")) as fp:
            recs = json.load(fp)[name]
        # cells with a non-trivial outcome are the informative ones
        good = [r for r in recs if 0 < r["ratio_noise"] < 1]
        pick = rng.choice(len(good), size=40, replace=False)
        sample[name] = [
            {k: good[i][k] for k in ("n_points", "radius_cutoff", "cnn_cutoff", "member_cutoff", "n_clusters",
                                     "ratio_largest", "ratio_noise")}
            for i in sorted(pick)
        ]
    with open(os.path.join(HERE, "records_sample.json"), "w") as fp:
        json.dump(sample, fp, indent=0)
    print("records sampled:", {k: len(v) for k, v in sample.items()})


if __name__ == "__main__":
    main()
