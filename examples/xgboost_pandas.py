#!/usr/bin/env python
"""sparkdl.xgboost on a pandas DataFrame: validation indicator, weights, early stopping, persistence, warm start.

    python examples/xgboost_pandas.py [--num-workers 2] [--use-gpu]
"""
import argparse
import tempfile

import numpy as np
import pandas as pd

from sparkdl.xgboost import XgboostClassifier, XgboostClassifierModel


def main(num_workers=1, use_gpu=False, rows=4000):
    rng = np.random.RandomState(0)
    X = rng.randn(rows, 12).astype(np.float32)
    y = ((X[:, 0] * X[:, 1] + np.sin(2 * X[:, 2]) + 0.2 * rng.randn(rows)) > 0).astype(np.float32)
    X[rng.rand(*X.shape) < 0.05] = 0.0                      # treat 0.0 as "missing", as the reference recommends
    df = pd.DataFrame({"features": list(X), "label": y, "isVal": rng.rand(rows) < 0.2, "weight": rng.uniform(0.5, 2, rows)})
    clf = XgboostClassifier(n_estimators=60, max_depth=4, learning_rate=0.2, missing=0.0, validationIndicatorCol="isVal",
                            weightCol="weight", early_stopping_rounds=5, eval_metric="logloss", num_workers=num_workers,
                            use_gpu=use_gpu)
    model = clf.fit(df)
    out = model.transform(df)
    acc = float((out["prediction"] == df["label"]).mean())
    with tempfile.TemporaryDirectory() as d:
        model.write().overwrite().save(d + "/model")
        again = XgboostClassifierModel.load(d + "/model")
        same = bool(np.allclose(np.stack(again.transform(df)["probability"].to_list()), np.stack(out["probability"].to_list())))
    more = XgboostClassifier(n_estimators=5, max_depth=4, xgb_model=model.get_booster()).fit(df[["features", "label"]])
    return {"accuracy": acc, "best_iteration": model.get_booster().best_iteration, "reloaded_identical": same,
            "rounds_after_warm_start": more.get_booster().num_boosted_rounds}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-workers", type=int, default=1)
    ap.add_argument("--use-gpu", action="store_true")
    a = ap.parse_args()
    print(main(a.num_workers, a.use_gpu))
