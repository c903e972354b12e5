"""The Spark code paths against ``tests/fakes/pyspark`` (a JVM-free stand-in with the same API surface): ``np > 0`` as a
barrier stage (ref runner_base.py:54-61), Spark-native distributed xgboost (ref xgboost.py:58-64), and the
``HAVE_PYSPARK`` branches of the estimator (pyspark base classes, MLWriter-style persistence).  Runs in a child process
so that the fake never leaks into the other tests."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, json
import numpy as np, pandas as pd
from pyspark.sql import SparkSession
import pyspark
assert pyspark.__version__ == "0.0-fake"
spark = SparkSession.builder.master("local[4]").getOrCreate()
from spark_deep_learning_b200.xgboost import params
assert params.HAVE_PYSPARK
from sparkdl import HorovodRunner
from sparkdl.xgboost import XgboostRegressor, XgboostClassifier, XgboostRegressorModel
res = {}

def main(scale):
    import torch
    import sparkdl.hvd as hvd
    from sparkdl.horovod import log_to_driver
    hvd.init()
    log_to_driver("hello from rank %d" % hvd.rank())
    s = hvd.allreduce(torch.full((5,), float(hvd.rank() + 1) * scale), op=hvd.Sum)
    return {"rank": hvd.rank(), "size": hvd.size(), "sum": float(s[0]), "store": os.environ["SPARKDL_STORE"]}

out = HorovodRunner(np=3).run(main, scale=2.0)
res["gang"] = out
try:
    HorovodRunner(np=8).run(main, scale=1.0)
    res["oversubscribed"] = "ran"
except RuntimeError as e:
    res["oversubscribed"] = str(e)

rng = np.random.RandomState(0)
X = rng.randn(3000, 6).astype(np.float32)
y = (2 * X[:, 0] + np.sin(3 * X[:, 1]) + 0.1 * rng.randn(3000)).astype(np.float32)
pdf = pd.DataFrame({"features": list(X), "label": y, "isVal": rng.rand(3000) < 0.2})
sdf = spark.createDataFrame(pdf)
kw = dict(n_estimators=8, max_depth=4, learning_rate=0.3, validationIndicatorCol="isVal")
single = XgboostRegressor(**kw).fit(pdf).transform(pdf)["prediction"].to_numpy()
model = XgboostRegressor(num_workers=2, force_repartition=True, **kw).fit(sdf)
dist = model.transform(pdf)["prediction"].to_numpy()
res["xgb_workers"] = model.get_booster().attributes.get("workers")
res["xgb_rmse_single"] = float(np.sqrt(np.mean((single - y) ** 2)))
res["xgb_rmse_dist"] = float(np.sqrt(np.mean((dist - y) ** 2)))
# transform of a Spark DataFrame returns a Spark DataFrame
tdf = model.transform(sdf)
res["transform_type"] = type(tdf).__name__
res["transform_cols"] = sorted(tdf.columns)
# MLWriter-style persistence (HAVE_PYSPARK branch)
path = os.path.join(sys.argv[1], "model")
model.write().overwrite().save(path)
m2 = XgboostRegressorModel.read().load(path)
res["reload_equal"] = bool(np.allclose(m2.transform(pdf)["prediction"].to_numpy(), dist))
yc = (y > 0).astype(np.float32)
cdf = spark.createDataFrame(pd.DataFrame({"features": list(X), "label": yc}))
cm = XgboostClassifier(num_workers=2, n_estimators=5, max_depth=3).fit(cdf)
pc = cm.transform(pd.DataFrame({"features": list(X), "label": yc}))
res["clf_acc"] = float((pc["prediction"].to_numpy() == yc).mean())
print("RESULT " + json.dumps(res))
'''


def test_spark_paths_on_the_fake(tmp_path):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests", "fakes"), ROOT, env.get("PYTHONPATH", "")])
    env["SPARKDL_NUM_GPUS"] = "0"
    env["SPARKDL_FORCE_CPU"] = "1"
    out = subprocess.run([sys.executable, "-c", CHILD, str(tmp_path)], capture_output=True, text=True, timeout=600, env=env)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, out.stdout[-3000:] + out.stderr[-3000:]
    import json
    res = json.loads(line[0][7:])
    g = res["gang"]
    assert g["rank"] == 0 and g["size"] == 3 and g["sum"] == 12.0 and "sparkdl-spark-" in g["store"]
    assert "hello from rank" in out.stdout                      # log_to_driver reaches the driver's stdout
    assert "hello from rank 2" in out.stderr                    # ... and rank 0's (the first executor's) stderr
    assert "task slots" in res["oversubscribed"]                # np > slots: "the job will fail"
    assert res["xgb_workers"] == 2
    assert res["xgb_rmse_dist"] < 1.3 * res["xgb_rmse_single"] + 0.05
    assert res["transform_type"] == "DataFrame" and "prediction" in res["transform_cols"]
    assert res["reload_equal"] and res["clf_acc"] > 0.9
