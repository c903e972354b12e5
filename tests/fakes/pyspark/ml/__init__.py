"""pyspark.ml of the fake: the repo's own dependency-free Param machinery under pyspark's names (a private second
instance of ``params.py`` loaded with pyspark hidden, so that its shim branch runs)."""
import importlib.util
import sys

_hidden = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "pyspark.ml" or k.startswith("pyspark.ml.")}
sys.modules["pyspark.ml"] = None          # `from pyspark.ml import ...` inside params.py -> ImportError -> shim classes
try:
    _path = importlib.util.find_spec("spark_deep_learning_b200.xgboost").submodule_search_locations[0] + "/params.py"
    _spec = importlib.util.spec_from_file_location("fake_pyspark_ml_impl", _path)
    _p = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(_p)
finally:
    del sys.modules["pyspark.ml"]
    sys.modules.update(_hidden)

Estimator, Model, Transformer = _p.Estimator, _p.Model, _p.Transformer
