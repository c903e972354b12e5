"""Histogram gradient-boosted trees, written from scratch (xgboost itself is not installable here).

This is the engine under ``sparkdl.xgboost`` (reference contract:
``/root/reference/sparkdl/xgboost/xgboost.py:165-331``): squared-error regression, binary logistic
and multi-class soft-prob objectives, ``missing`` handling with a learned default direction,
instance weights, L2/L1 regularisation, ``gamma`` / ``min_child_weight`` / ``max_depth``,
row and column subsampling, evaluation set + early stopping, callbacks, warm start
(``xgb_model``), and data-parallel training where each worker owns a row shard and the per-level
histograms are summed across workers - by the fused sm_100a
``gbdt_hist_allreduce`` kernel on GPUs, or through the shm engine on CPU workers.

Gradients are quantised to int32 fixed point each round, so histograms are exact int64 sums:
training is bit-reproducible and independent of the worker count.
"""
import ctypes
import json
import math
import os

import numpy as np
import torch

MISSING_BIN = 255
MAX_BINS = 256


# ------------------------------------------------------------------------------------------------
# objectives / metrics
# ------------------------------------------------------------------------------------------------
def _sigmoid(x):
    return torch.sigmoid(x)


_LOGISTIC = ("binary:logistic", "reg:logistic", "binary:logitraw")
_LOG_LINK = ("count:poisson", "reg:gamma", "reg:tweedie")


class _Objective(object):
    """Gradient / hessian, link function and default metric of the supported xgboost objectives."""

    def __init__(self, name, num_class=1, max_delta_step=0.0, tweedie_variance_power=1.5):
        self.name = name
        self.num_class = num_class if name.startswith("multi:") else 1
        self.poisson_mds = max_delta_step if max_delta_step > 0 else 0.7      # xgboost's default for count:poisson
        self.rho = float(tweedie_variance_power)
        if name == "reg:tweedie" and not 1.0 <= self.rho < 2.0:
            raise ValueError("tweedie_variance_power must be in [1, 2)")

    @property
    def n_groups(self):
        return self.num_class

    def base_margin(self, base_score):
        if self.name in _LOGISTIC:
            p = min(max(base_score, 1e-6), 1 - 1e-6)
            return math.log(p / (1 - p))
        if self.name in _LOG_LINK:
            return math.log(max(base_score, 1e-16))
        return base_score

    def grad_hess(self, margin, y, w):
        if self.name in ("reg:squarederror", "reg:linear"):
            g, h = margin[:, 0] - y, torch.ones_like(y)
            return (g * w).unsqueeze(1), (h * w).unsqueeze(1)
        if self.name in _LOGISTIC:
            p = _sigmoid(margin[:, 0])
            return ((p - y) * w).unsqueeze(1), (torch.clamp(p * (1 - p), min=1e-16) * w).unsqueeze(1)
        m = margin[:, 0]
        if self.name == "reg:squaredlogerror":
            if bool((y <= -1).any()):
                raise ValueError("reg:squaredlogerror needs labels > -1")
            p = torch.clamp(m, min=-1 + 1e-6)
            d = torch.log1p(p) - torch.log1p(y)
            g, h = d / (p + 1), torch.clamp((1 - d) / (p + 1) ** 2, min=1e-6)
            return (g * w).unsqueeze(1), (h * w).unsqueeze(1)
        if self.name == "reg:pseudohubererror":
            z = m - y
            sc = 1 + z * z
            return (z / torch.sqrt(sc) * w).unsqueeze(1), (w / (sc * torch.sqrt(sc))).unsqueeze(1)
        if self.name == "count:poisson":
            if bool((y < 0).any()):
                raise ValueError("count:poisson needs non-negative labels")
            return ((torch.exp(m) - y) * w).unsqueeze(1), (torch.exp(m + self.poisson_mds) * w).unsqueeze(1)
        if self.name == "reg:gamma":
            if bool((y <= 0).any()):
                raise ValueError("reg:gamma needs positive labels")
            r = y / torch.exp(m)
            return ((1 - r) * w).unsqueeze(1), (r * w).unsqueeze(1)
        if self.name == "reg:tweedie":
            if bool((y < 0).any()):
                raise ValueError("reg:tweedie needs non-negative labels")
            a, b = torch.exp((1 - self.rho) * m), torch.exp((2 - self.rho) * m)
            g = -y * a + b
            h = -y * (1 - self.rho) * a + (2 - self.rho) * b
            return (g * w).unsqueeze(1), (h * w).unsqueeze(1)
        if self.name == "binary:hinge":
            t = 2 * y - 1
            act = m * t < 1
            g = torch.where(act, -t, torch.zeros_like(m))
            h = torch.where(act, torch.ones_like(m), torch.full_like(m, 1e-16))
            return (g * w).unsqueeze(1), (h * w).unsqueeze(1)
        if self.name in ("multi:softprob", "multi:softmax"):
            p = torch.softmax(margin, dim=1)
            onehot = torch.nn.functional.one_hot(y.long(), self.num_class).to(p.dtype)
            return (p - onehot) * w.unsqueeze(1), torch.clamp(2.0 * p * (1 - p), min=1e-16) * w.unsqueeze(1)
        raise ValueError("unsupported objective %r" % self.name)

    def transform(self, margin):
        if self.name in ("binary:logistic", "reg:logistic"):
            return _sigmoid(margin[:, 0])
        if self.name in _LOG_LINK:
            return torch.exp(margin[:, 0])
        if self.name == "binary:hinge":
            return (margin[:, 0] > 0).to(margin.dtype)
        if self.name in ("multi:softprob", "multi:softmax"):
            return torch.softmax(margin, dim=1)
        return margin[:, 0]

    def default_metric(self):
        return {"reg:squarederror": "rmse", "reg:linear": "rmse", "binary:logistic": "logloss", "binary:logitraw": "logloss",
                "reg:logistic": "rmse", "multi:softprob": "mlogloss", "multi:softmax": "mlogloss",
                "reg:squaredlogerror": "rmsle", "reg:pseudohubererror": "mphe", "count:poisson": "poisson-nloglik",
                "reg:gamma": "gamma-nloglik", "reg:tweedie": "tweedie-nloglik@%g" % self.rho,
                "binary:hinge": "error"}[self.name]


def _metric_sums(name, obj, margin, y, w):
    """Return (weighted sum, weight) so workers can be combined with one allreduce."""
    eps = 1e-15
    if name == "rmse":
        pred = obj.transform(margin)
        return float(((pred - y) ** 2 * w).sum()), float(w.sum())
    if name == "mae":
        pred = obj.transform(margin)
        return float(((pred - y).abs() * w).sum()), float(w.sum())
    if name == "logloss":
        p = torch.clamp(_sigmoid(margin[:, 0]), eps, 1 - eps)
        return float((-(y * torch.log(p) + (1 - y) * torch.log(1 - p)) * w).sum()), float(w.sum())
    if name == "error" or name.startswith("error@"):
        thr = float(name.split("@", 1)[1]) if "@" in name else 0.5
        p = obj.transform(margin) if obj.name == "binary:hinge" else _sigmoid(margin[:, 0])
        return float((((p > thr).to(y.dtype) != y).to(w.dtype) * w).sum()), float(w.sum())
    if name == "rmsle":
        pred = torch.clamp(obj.transform(margin), min=-1 + 1e-6)
        return float(((torch.log1p(pred) - torch.log1p(y)) ** 2 * w).sum()), float(w.sum())
    if name == "mape":
        pred = obj.transform(margin)
        return float((((pred - y) / y).abs() * w).sum()), float(w.sum())
    if name == "mphe":
        z = obj.transform(margin) - y
        return float(((torch.sqrt(1 + z * z) - 1) * w).sum()), float(w.sum())
    if name == "poisson-nloglik":
        pred = torch.clamp(obj.transform(margin), min=eps)
        return float(((pred - y * torch.log(pred) + torch.lgamma(y + 1)) * w).sum()), float(w.sum())
    if name == "gamma-nloglik":                       # unit dispersion, as in xgboost
        pred = torch.clamp(obj.transform(margin), min=eps)
        return float(((y / pred + torch.log(pred)) * w).sum()), float(w.sum())
    if name == "gamma-deviance":
        pred = torch.clamp(obj.transform(margin), min=eps)
        r = torch.clamp(y, min=eps) / pred
        return float((2 * (r - torch.log(r) - 1) * w).sum()), float(w.sum())
    if name.startswith("tweedie-nloglik"):
        rho = float(name.split("@", 1)[1]) if "@" in name else 1.5
        pred = torch.clamp(obj.transform(margin), min=eps)
        a = y * torch.pow(pred, 1 - rho) / (1 - rho)
        b = torch.pow(pred, 2 - rho) / (2 - rho)
        return float(((-a + b) * w).sum()), float(w.sum())
    if name == "mlogloss":
        p = torch.clamp(torch.softmax(margin, 1), eps, 1.0)
        return float((-torch.log(p.gather(1, y.long().unsqueeze(1)))[:, 0] * w).sum()), float(w.sum())
    if name == "merror":
        return float(((margin.argmax(1) != y.long()).to(w.dtype) * w).sum()), float(w.sum())
    raise ValueError("unsupported eval_metric %r" % name)


def _finish_metric(name, s, n):
    v = s / max(n, 1e-12)
    return math.sqrt(v) if name in ("rmse", "rmsle") else v


def _auc(score, y, w):
    order = torch.argsort(score, descending=True, stable=True)
    y, w = y[order], w[order]
    tp = torch.cumsum(y * w, 0)
    fp = torch.cumsum((1 - y) * w, 0)
    if tp[-1] <= 0 or fp[-1] <= 0:
        return 0.5
    tpr = torch.cat([tp.new_zeros(1), tp / tp[-1]])
    fpr = torch.cat([fp.new_zeros(1), fp / fp[-1]])
    return float(torch.trapz(tpr, fpr))


# ------------------------------------------------------------------------------------------------
# trees
# ------------------------------------------------------------------------------------------------
class Tree(object):
    """Flat arrays: left/right child index (-1 at leaves), split feature, threshold (x < thr goes
    left), default direction for missing values, leaf value, cover, gain."""

    def __init__(self):
        self.feature, self.threshold, self.left, self.right = [], [], [], []
        self.default_left, self.value, self.cover, self.gain = [], [], [], []

    def add_node(self):
        for a, v in ((self.feature, -1), (self.threshold, 0.0), (self.left, -1), (self.right, -1),
                     (self.default_left, 1), (self.value, 0.0), (self.cover, 0.0), (self.gain, 0.0)):
            a.append(v)
        return len(self.left) - 1

    def to_dict(self):
        return {k: getattr(self, k) for k in ("feature", "threshold", "left", "right", "default_left",
                                              "value", "cover", "gain")}

    @staticmethod
    def from_dict(d):
        t = Tree()
        for k, v in d.items():
            setattr(t, k, list(v))
        return t

    def depth(self):
        def rec(i):
            return 0 if self.left[i] < 0 else 1 + max(rec(self.left[i]), rec(self.right[i]))
        return rec(0) if self.left else 0


class Booster(object):
    """The trained ensemble - what ``XgboostRegressorModel.get_booster()`` returns
    (ref ``xgboost.py:130-133``)."""

    def __init__(self, objective="reg:squarederror", num_class=1, base_score=0.5, missing=float("nan"),
                 n_features=0):
        self.objective = objective
        self.num_class = num_class
        self.base_score = base_score
        self.missing = missing
        self.n_features = n_features
        self.trees = []          # list[Tree]
        self.tree_group = []     # output group (class) of each tree
        self.best_iteration = None
        self.best_score = None
        self.attributes = {}
        self._packed = None

    # -- persistence ----------------------------------------------------------------------------------
    def to_json(self):
        miss = None if (isinstance(self.missing, float) and math.isnan(self.missing)) else self.missing
        return json.dumps({"format": "sparkdl-b200-gbdt-1", "objective": self.objective, "num_class": self.num_class,
                           "base_score": self.base_score, "missing": miss, "n_features": self.n_features,
                           "best_iteration": self.best_iteration, "best_score": self.best_score,
                           "attributes": self.attributes, "tree_group": self.tree_group,
                           "trees": [t.to_dict() for t in self.trees]})

    @staticmethod
    def from_json(s):
        d = json.loads(s)
        b = Booster(d["objective"], d["num_class"], d["base_score"],
                    float("nan") if d["missing"] is None else d["missing"], d["n_features"])
        b.trees = [Tree.from_dict(t) for t in d["trees"]]
        b.tree_group = list(d["tree_group"])
        b.best_iteration, b.best_score = d.get("best_iteration"), d.get("best_score")
        b.attributes = d.get("attributes", {})
        return b

    # xgboost's own JSON model schema (what ``xgboost.Booster.save_model("m.json")`` writes and ``load_model`` reads):
    # ``get_booster()`` of the reference returns an ``xgboost.core.Booster`` (ref ``xgboost.py:130-133``); xgboost is not
    # installable here, so interoperability goes through the file format - a model saved with ``save_model("x.json")``
    # opens in stock xgboost, and a gbtree model written by stock xgboost loads here.
    _OBJ_PARAM = {"reg:": ("reg_loss_param", {"scale_pos_weight": "1"}), "binary:": ("reg_loss_param", {"scale_pos_weight": "1"}),
                  "count:poisson": ("poisson_regression_param", {"max_delta_step": "0.7"}),
                  "reg:tweedie": ("tweedie_regression_param", {"tweedie_variance_power": "1.5"}),
                  "reg:pseudohubererror": ("pseudo_huber_param", {"huber_slope": "1"})}

    def to_xgboost_json(self):
        trees = []
        for tid, t in enumerate(self.trees):
            n = len(t.left)
            parents = [2147483647] * n
            for i in range(n):
                if t.left[i] >= 0:
                    parents[t.left[i]] = i
                    parents[t.right[i]] = i
            leaf = [t.left[i] < 0 for i in range(n)]
            trees.append({
                "base_weights": [float(v) for v in t.value], "categories": [], "categories_nodes": [],
                "categories_segments": [], "categories_sizes": [],
                "default_left": [int(bool(d)) for d in t.default_left], "id": tid,
                "left_children": [int(v) for v in t.left], "loss_changes": [float(g) for g in t.gain], "parents": parents,
                "right_children": [int(v) for v in t.right],
                # xgboost keeps the leaf value in split_conditions of a leaf
                "split_conditions": [float(t.value[i]) if leaf[i] else float(t.threshold[i]) for i in range(n)],
                "split_indices": [0 if leaf[i] else int(t.feature[i]) for i in range(n)], "split_type": [0] * n,
                "sum_hessian": [float(c) for c in t.cover],
                "tree_param": {"num_deleted": "0", "num_feature": str(self.n_features), "num_nodes": str(n),
                               "size_leaf_vector": "1"}})
        npt = max(1, int(self.attributes.get("num_parallel_tree", 1)))
        per_round = max(1, self.n_groups) * npt
        rounds = len(self.trees) // per_round if per_round else 0
        multi = self.objective.startswith("multi:")
        if multi:
            obj = {"name": self.objective, "softmax_multiclass_param": {"num_class": str(self.num_class)}}
        else:
            key = next((k for k in sorted(self._OBJ_PARAM, key=len, reverse=True) if self.objective.startswith(k)), "reg:")
            pname, pval = self._OBJ_PARAM[key]
            obj = {"name": self.objective, pname: dict(pval)}
        attrs = {k: (v if isinstance(v, str) else json.dumps(v)) for k, v in self.attributes.items()}
        if self.best_iteration is not None:
            attrs["best_iteration"] = str(self.best_iteration)
        if self.best_score is not None:
            attrs["best_score"] = repr(float(self.best_score))
        if not (isinstance(self.missing, float) and math.isnan(self.missing)):
            attrs["sparkdl_missing"] = repr(float(self.missing))
        return json.dumps({"learner": {
            "attributes": attrs, "feature_names": [], "feature_types": [],
            "gradient_booster": {"model": {
                "gbtree_model_param": {"num_parallel_tree": str(npt), "num_trees": str(len(self.trees))},
                "iteration_indptr": [i * per_round for i in range(rounds + 1)],
                "tree_info": [int(g) for g in self.tree_group], "trees": trees}, "name": "gbtree"},
            "learner_model_param": {"base_score": "%.9E" % float(self.base_score), "boost_from_average": "1",
                                    "num_class": str(self.num_class if multi else 0),
                                    "num_feature": str(self.n_features), "num_target": "1"},
            "objective": obj}, "version": [2, 0, 3]})

    @staticmethod
    def from_xgboost_json(text):
        d = json.loads(text) if isinstance(text, str) else text
        L = d["learner"]
        gb = L["gradient_booster"]
        if gb.get("name", "gbtree") not in ("gbtree", "dart"):
            raise ValueError("only gbtree models can be loaded, got %r" % gb.get("name"))
        model = gb["model"] if "model" in gb else gb["gbtree"]["model"]
        lmp = L["learner_model_param"]
        objective = L["objective"]["name"]
        num_class = max(1, int(lmp.get("num_class", "0")))
        attrs = dict(L.get("attributes", {}))
        missing = float(attrs.pop("sparkdl_missing")) if "sparkdl_missing" in attrs else float("nan")
        b = Booster(objective, num_class, float(lmp.get("base_score", "0.5")), missing, int(lmp.get("num_feature", "0")))
        if "best_iteration" in attrs:
            b.best_iteration = int(attrs.pop("best_iteration"))
        if "best_score" in attrs:
            b.best_score = float(attrs.pop("best_score"))
        for k, v in attrs.items():
            try:
                b.attributes[k] = json.loads(v)
            except (ValueError, TypeError):
                b.attributes[k] = v
        npt = int(model.get("gbtree_model_param", {}).get("num_parallel_tree", "1"))
        if npt > 1:
            b.attributes["num_parallel_tree"] = npt
        for t in model["trees"]:
            if any(int(x) != 0 for x in t.get("split_type", [])):
                raise ValueError("categorical splits are not supported")
            tr = Tree()
            n = len(t["left_children"])
            for i in range(n):
                tr.add_node()
                leaf = t["left_children"][i] < 0
                tr.left[i], tr.right[i] = int(t["left_children"][i]), int(t["right_children"][i])
                tr.default_left[i] = int(t["default_left"][i])
                tr.feature[i] = -1 if leaf else int(t["split_indices"][i])
                tr.threshold[i] = 0.0 if leaf else float(t["split_conditions"][i])
                tr.value[i] = float(t["split_conditions"][i]) if leaf else float(t["base_weights"][i])
                tr.cover[i] = float(t["sum_hessian"][i])
                tr.gain[i] = float(t["loss_changes"][i])
            b.trees.append(tr)
        b.tree_group = [int(g) for g in model["tree_info"]]
        return b

    def save_model(self, path):
        """``*.json`` -> xgboost's JSON model schema (opens in stock xgboost); anything else -> the native format."""
        with open(path, "w") as f:
            f.write(self.to_xgboost_json() if str(path).endswith(".json") else self.to_json())

    @staticmethod
    def load_model(path):
        with open(path) as f:
            text = f.read()
        d = json.loads(text)
        return Booster.from_xgboost_json(d) if "learner" in d else Booster.from_json(text)

    def copy(self):
        return Booster.from_json(self.to_json())

    @property
    def trees_per_round(self):
        """Output groups x ``num_parallel_tree`` (a boosted random forest adds several trees per round)."""
        return max(1, self.n_groups) * max(1, int(self.attributes.get("num_parallel_tree", 1)))

    @property
    def num_boosted_rounds(self):
        return len(self.trees) // self.trees_per_round

    def get_dump(self):
        out = []
        for t in self.trees:
            lines = []

            def rec(i, d, t=t, lines=lines):
                pad = "\t" * d
                if t.left[i] < 0:
                    lines.append("%s%d:leaf=%.6g" % (pad, i, t.value[i]))
                else:
                    lines.append("%s%d:[f%d<%.6g] yes=%d,no=%d,missing=%d" % (
                        pad, i, t.feature[i], t.threshold[i], t.left[i], t.right[i],
                        t.left[i] if t.default_left[i] else t.right[i]))
                    rec(t.left[i], d + 1)
                    rec(t.right[i], d + 1)
            rec(0, 0)
            out.append("\n".join(lines))
        return out

    def get_score(self, fmap="", importance_type="weight"):
        """Feature importance as ``xgboost.Booster.get_score``: ``weight`` (number of splits on the feature), ``gain`` /
        ``cover`` (average loss reduction / average hessian mass of those splits), ``total_gain``, ``total_cover``.
        Features that are never used are absent.  ``fmap`` is accepted for signature compatibility."""
        if fmap not in ("", None) and importance_type == "weight" and fmap in ("weight", "gain", "cover", "total_gain", "total_cover"):
            importance_type = fmap                   # get_score("gain"): positional use of the pre-2.0 signature
        if importance_type not in ("weight", "gain", "cover", "total_gain", "total_cover"):
            raise ValueError("unknown importance_type %r" % (importance_type,))
        count, gain, cover = {}, {}, {}
        for t in self.trees:
            for i, f in enumerate(t.feature):
                if t.left[i] >= 0:
                    k = "f%d" % f
                    count[k] = count.get(k, 0) + 1
                    gain[k] = gain.get(k, 0.0) + float(t.gain[i])
                    cover[k] = cover.get(k, 0.0) + float(t.cover[i])
        if importance_type == "weight":
            return {k: float(v) for k, v in count.items()}
        if importance_type == "total_gain":
            return gain
        if importance_type == "total_cover":
            return cover
        src = gain if importance_type == "gain" else cover
        return {k: src[k] / count[k] for k in count}

    def feature_importances(self, importance_type="gain"):
        """Dense, normalised importance vector of length ``n_features`` (what ``XGBModel.feature_importances_`` holds)."""
        sc = self.get_score(importance_type=importance_type)
        n = int(self.n_features) if self.n_features else (max([int(k[1:]) for k in sc] + [-1]) + 1)
        v = np.zeros(n, dtype=np.float32)
        for k, x in sc.items():
            v[int(k[1:])] = x
        tot = float(v.sum())
        return v / tot if tot > 0 else v

    # -- inference ------------------------------------------------------------------------------------
    def _pack(self, device, ntree_limit=None):
        nt = len(self.trees) if not ntree_limit else min(len(self.trees), ntree_limit)
        key = (str(device), nt)
        if self._packed is not None and self._packed[0] == key:
            return self._packed[1]
        off, feat, thr, left, right, dl, val = [], [], [], [], [], [], []
        for t in self.trees[:nt]:
            off.append(len(feat))
            feat += [max(f, 0) for f in t.feature]
            thr += t.threshold
            left += t.left
            right += t.right
            dl += t.default_left
            val += t.value
        mk = lambda a, dt: torch.tensor(a if a else [0], dtype=dt, device=device)  # noqa: E731
        packed = {"feat": mk(feat, torch.int32), "thr": mk(thr, torch.float32), "left": mk(left if left else [-1], torch.int32),
                  "right": mk(right if right else [-1], torch.int32), "dleft": mk(dl, torch.uint8), "value": mk(val, torch.float32),
                  "off": mk(off, torch.int32), "group": mk(self.tree_group[:nt], torch.int32), "n_trees": nt}
        self._packed = (key, packed)
        return packed

    @property
    def n_groups(self):
        return self.num_class if self.objective.startswith("multi:") else 1

    def predict_margin(self, X, base_margin=None, ntree_limit=None, zero_base=False):
        """X: float32 [n, F] tensor (CPU or CUDA).  Returns margins [n, n_groups]."""
        X = torch.as_tensor(X, dtype=torch.float32)
        n, F = X.shape
        obj = _Objective(self.objective, self.num_class)
        base = 0.0 if zero_base else float(obj.base_margin(self.base_score))
        out = torch.full((n, self.n_groups), base, dtype=torch.float32, device=X.device)
        if base_margin is not None:
            out = torch.tensor(np.array(base_margin, dtype=np.float32), device=X.device).reshape(n, -1)
        if not self.trees:
            return out
        p = self._pack(X.device, ntree_limit)
        miss_nan = isinstance(self.missing, float) and math.isnan(self.missing)
        if X.is_cuda:
            lib = _cuda_lib()
            Xc = X.contiguous()
            rc = lib.sdlc_gbdt_predict(
                _vp(Xc), n, F, 0.0 if miss_nan else float(self.missing), 1 if miss_nan else 0, _vp(p["feat"]), _vp(p["thr"]),
                _vp(p["left"]), _vp(p["right"]), _vp(p["dleft"]), _vp(p["value"]), _vp(p["off"]), _vp(p["group"]),
                p["n_trees"], self.n_groups, _vp(out), torch.cuda.current_stream(X.device).cuda_stream)
            if rc != 0:
                raise RuntimeError("sdlc_gbdt_predict failed")
            from spark_deep_learning_b200.utils import launch_counter
            launch_counter.add(1)
            return out
        # CPU: vectorised level-synchronous traversal, one tree at a time
        miss = torch.isnan(X) if miss_nan else ((X == float(self.missing)) | torch.isnan(X))
        rows = torch.arange(n)
        for ti in range(p["n_trees"]):
            base = int(p["off"][ti])
            node = torch.zeros(n, dtype=torch.long)
            while True:
                l = p["left"][base + node].long()
                active = l >= 0
                if not bool(active.any()):
                    break
                f = p["feat"][base + node].long()
                v = X[rows, f]
                m = miss[rows, f]
                go_left = torch.where(m, p["dleft"][base + node] != 0, v < p["thr"][base + node])
                nxt = torch.where(go_left, l, p["right"][base + node].long())
                node = torch.where(active, nxt, node)
            out[:, int(p["group"][ti])] += p["value"][base + node]
        return out

    def predict(self, X, output_margin=False, base_margin=None, ntree_limit=None):
        m = self.predict_margin(X, base_margin, ntree_limit)
        if output_margin:
            return m if self.n_groups > 1 else m[:, 0]
        return _Objective(self.objective, self.num_class).transform(m)


def _vp(t):
    return ctypes.c_void_p(t.data_ptr())


_LIB_SIG = False


def _cuda_lib():
    global _LIB_SIG
    from spark_deep_learning_b200.ops import comm
    lib = comm.load()
    if not _LIB_SIG:
        vp, i32, i64, u64, f32, u32 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_ulonglong,
                                        ctypes.c_float, ctypes.c_uint)
        lib.sdlc_gbdt_hist.restype = i32
        lib.sdlc_gbdt_hist.argtypes = [vp, vp, vp, vp, i64, i32, u64]
        lib.sdlc_gbdt_hist_smem.restype = i32
        lib.sdlc_gbdt_hist_smem.argtypes = [vp, i64, vp, vp, vp, i64, i32, i32, u64]
        lib.sdlc_gbdt_hist_allreduce.restype = i32
        lib.sdlc_gbdt_hist_allreduce.argtypes = [vp, i32, i32, vp, vp, i64, vp, vp, i64, i32, i32, i64, vp, u32, u64]
        lib.sdlc_gbdt_predict.restype = i32
        lib.sdlc_gbdt_predict.argtypes = [vp, i64, i32, f32, i32] + [vp] * 8 + [i32, i32, vp, u64]
        _LIB_SIG = True
    return lib


# ------------------------------------------------------------------------------------------------
# training
# ------------------------------------------------------------------------------------------------
class _Dist(object):
    """Data-parallel context: identity when single worker, hvd otherwise."""

    def __init__(self):
        self.hvd = None
        try:
            import sys
            hvd = sys.modules.get("spark_deep_learning_b200.parallel.hvd")
            if hvd is not None and hvd.is_initialized() and hvd.size() > 1:
                self.hvd = hvd
        except Exception:  # pylint: disable=broad-except
            self.hvd = None
        self.size = self.hvd.size() if self.hvd else 1
        self.rank = self.hvd.rank() if self.hvd else 0
        self._n = 0

    def _name(self, tag):
        self._n += 1
        return "gbdt.%s.%d" % (tag, self._n)

    def allreduce_host(self, values, op="sum"):
        if not self.hvd:
            return values
        t = torch.tensor(values, dtype=torch.float64)
        op_ = {"sum": self.hvd.Sum, "max": self.hvd.Max, "min": self.hvd.Min}[op]
        return self.hvd.allreduce(t, op=op_, name=self._name("host")).tolist()

    def allgather(self, t):
        if not self.hvd:
            return t
        return self.hvd.allgather(t.cpu(), name=self._name("gather")).to(t.device)

    def allreduce_tensor(self, t):
        """int64 histogram sum (generic path; the fused GPU kernel bypasses this)."""
        if not self.hvd:
            return t
        return self.hvd.allreduce(t, op=self.hvd.Sum, name=self._name("hist"))


def compute_cuts(X, missing, max_bin, dist, sample_rows=200000, seed=0):
    """Per-feature quantile cut points from a (distributed) row sample.  Returns list of 1-D float32
    tensors, at most ``max_bin - 1`` cuts each (bin 255 is reserved for missing values)."""
    n, F = X.shape
    g = torch.Generator(device="cpu").manual_seed(seed + dist.rank)
    k = min(n, max(1, sample_rows // dist.size))
    idx = torch.randperm(n, generator=g)[:k].to(X.device) if k < n else torch.arange(n, device=X.device)
    S = X[idx]
    if dist.size > 1:
        S = dist.allgather(S.contiguous())
    miss_nan = isinstance(missing, float) and math.isnan(missing)
    # c cuts make c + 1 real bins 0..c; bin 255 is reserved for missing values, so at most 254 cuts
    n_cuts = min(max_bin, MAX_BINS - 1) - 1
    cuts = []
    qs = torch.linspace(0, 1, n_cuts + 2, device=S.device)[1:-1]
    for f in range(F):
        col = S[:, f]
        col = col[~torch.isnan(col)] if miss_nan else col[(col != missing) & ~torch.isnan(col)]
        if col.numel() == 0:
            cuts.append(torch.zeros(0, dtype=torch.float32, device=X.device))
            continue
        u = torch.unique(col)
        if u.numel() <= n_cuts:
            c = (u[:-1] + u[1:]) / 2 if u.numel() > 1 else u.new_zeros(0)
        else:
            c = torch.unique(torch.quantile(col.float(), qs.float()))
        cuts.append(c.float().contiguous())
    return cuts


def bin_features(X, cuts, missing):
    n, F = X.shape
    miss_nan = isinstance(missing, float) and math.isnan(missing)
    B = torch.empty((n, F), dtype=torch.uint8, device=X.device)
    for f in range(F):
        col = X[:, f].contiguous()
        b = torch.searchsorted(cuts[f], col, right=True) if cuts[f].numel() else torch.zeros(n, dtype=torch.long, device=X.device)
        m = torch.isnan(col) if miss_nan else ((col == missing) | torch.isnan(col))
        b = b.clamp_(max=MISSING_BIN - 1)   # a real value never lands in the missing bin
        B[:, f] = torch.where(m, torch.full_like(b, MISSING_BIN), b).to(torch.uint8)
    return B.contiguous()


_CORE_HIST = {"fn": False}


def _core_hist_fn():
    """``sdl_gbdt_hist_cpu`` of libsparkdl_core.so, or None (``SPARKDL_GBDT_NATIVE_CPU=0`` / stale build)."""
    if _CORE_HIST["fn"] is False:
        fn = None
        if os.environ.get("SPARKDL_GBDT_NATIVE_CPU", "1") != "0":
            try:
                from spark_deep_learning_b200.parallel import core
                fn = getattr(core.load(), "sdl_gbdt_hist_cpu", None)
            except (OSError, RuntimeError):
                fn = None
        _CORE_HIST["fn"] = fn
    return _CORE_HIST["fn"]


class _HistBuilder(object):
    """Per-level histogram construction + cross-worker sum."""

    def __init__(self, bins, dist, use_fused=True):
        self.bins = bins
        self.n, self.F = bins.shape
        self.dist = dist
        self.cuda = bins.is_cuda
        self.sym = None
        self.comm = None
        self.grid_ctr = None
        self.grid_epoch = 0
        self.launches = 0
        # feature-major copy of the bins for the shared-memory privatised build (coalesced per-feature loads)
        self.binsT, self.pitch = None, 0
        if self.cuda and os.environ.get("SPARKDL_GBDT_SMEM_HIST", "1") != "0":
            self.pitch = (self.n + 15) // 16 * 16
            self.binsT = torch.zeros(self.F, self.pitch, dtype=torch.uint8, device=bins.device)
            self.binsT[:, :self.n] = bins.t()
        if self.cuda and dist.size > 1 and use_fused:
            self.comm = dist.hvd.comm_backend()   # None on a CPU-only gang -> generic allreduce path

    def _sym_hist(self, nelem):
        nbytes = nelem * 8
        if self.sym is None or self.sym.nbytes < nbytes:
            self.sym = self.comm.alloc_symmetric(max(nbytes, 1 << 20), multicast=False)
            self.grid_ctr = torch.zeros(1, dtype=torch.int32, device=self.bins.device)
            self.grid_epoch = 0
        return self.sym

    def build(self, gh_q, pos, n_nodes):
        """gh_q: int32 [n,2]; pos: int32 [n] (node slot or -1).  Returns int64 [n_nodes,F,256,2]."""
        nelem = n_nodes * self.F * MAX_BINS * 2
        if self.cuda:
            lib = _cuda_lib()
            stream = torch.cuda.current_stream(self.bins.device).cuda_stream
            if self.comm is not None:
                sym = self._sym_hist(nelem)
                hist = sym.tensor(torch.int64, nelem)
                hist.zero_()
                self.grid_epoch += 1
                rc = lib.sdlc_gbdt_hist_allreduce(self.comm._h, 2, sym.id, _vp(self.bins),
                                                  _vp(self.binsT) if self.binsT is not None else None, self.pitch, _vp(gh_q), _vp(pos),
                                                  self.n, self.F, n_nodes, nelem, _vp(self.grid_ctr), self.grid_epoch, stream)
                if rc < 0:
                    raise RuntimeError("gbdt_hist_allreduce failed: %s" % self.comm._err())
                # the device-local barrier counts `rc` CTAs per launch
                self._blocks = rc
                self.launches += 1
                return hist.view(n_nodes, self.F, MAX_BINS, 2).clone()
            hist = torch.zeros(nelem, dtype=torch.int64, device=self.bins.device)
            rc = 1
            if self.binsT is not None:
                rc = lib.sdlc_gbdt_hist_smem(_vp(self.binsT), self.pitch, _vp(gh_q), _vp(pos), _vp(hist), self.n, self.F, n_nodes, stream)
                if rc < 0:
                    raise RuntimeError("sdlc_gbdt_hist_smem failed")
            if rc == 1 and lib.sdlc_gbdt_hist(_vp(self.bins), _vp(gh_q), _vp(pos), _vp(hist), self.n, self.F, stream) != 0:
                raise RuntimeError("sdlc_gbdt_hist failed")
            self.launches += 1
            from spark_deep_learning_b200.utils import launch_counter
            launch_counter.add(1)
            hist = self.dist.allreduce_tensor(hist)
            return hist.view(n_nodes, self.F, MAX_BINS, 2)
        # CPU: native feature-parallel builder (libsparkdl_core), exact integers
        native = _core_hist_fn()
        if native is not None and n_nodes * self.F * MAX_BINS * 2 < (1 << 31):
            hist = torch.zeros(nelem, dtype=torch.int64)
            b, q, ps = self.bins.contiguous(), gh_q.contiguous(), pos.contiguous()
            if native(b.data_ptr(), q.data_ptr(), ps.data_ptr(), hist.data_ptr(), self.n, self.F, n_nodes, 0) != 0:
                raise RuntimeError("sdl_gbdt_hist_cpu rejected its arguments")
            hist = self.dist.allreduce_tensor(hist)
            return hist.view(n_nodes, self.F, MAX_BINS, 2)
        # fallback: torch index_add_
        act = pos >= 0
        rows = torch.nonzero(act)[:, 0]
        hist = torch.zeros(nelem, dtype=torch.int64)
        if rows.numel():
            b = self.bins[rows].long()
            base = (pos[rows].long().unsqueeze(1) * self.F + torch.arange(self.F).unsqueeze(0)) * MAX_BINS + b
            idx = (base * 2).reshape(-1)
            g = gh_q[rows, 0].long().unsqueeze(1).expand(-1, self.F).reshape(-1)
            h = gh_q[rows, 1].long().unsqueeze(1).expand(-1, self.F).reshape(-1)
            hist.index_add_(0, idx, g)
            hist.index_add_(0, idx + 1, h)
        hist = self.dist.allreduce_tensor(hist)
        return hist.view(n_nodes, self.F, MAX_BINS, 2)


IGNORED_PARAMS = ("booster", "n_jobs", "tree_method")      # importance_type is honoured by feature_importances_

DEFAULTS = dict(scale_pos_weight=1.0, n_estimators=100, max_depth=6, learning_rate=0.3, gamma=0.0, min_child_weight=1.0, subsample=1.0,
                colsample_bytree=1.0, reg_lambda=1.0, reg_alpha=0.0, max_bin=256, base_score=0.5, random_state=0,
                objective=None, num_class=None, early_stopping_rounds=None, eval_metric=None, verbose=False,
                max_delta_step=0.0, tree_method="hist", colsample_bylevel=1.0, colsample_bynode=1.0,
                monotone_constraints=None, tweedie_variance_power=1.5, interaction_constraints=None, num_parallel_tree=1)
_ALIASES = {"eta": "learning_rate", "lambda": "reg_lambda", "alpha": "reg_alpha", "seed": "random_state",
            "min_split_loss": "gamma", "num_boost_round": "n_estimators", "verbosity": "verbose"}


def normalize_params(params):
    p = dict(DEFAULTS)
    for k, v in params.items():
        k = _ALIASES.get(k, k)
        if v is not None or k in ("objective", "eval_metric", "early_stopping_rounds", "num_class"):
            p[k] = v
    return p


def parse_monotone(spec, F, feature_names=None):
    """``"(1,0,-1)"`` / sequence / ``{feature index or name: c}`` -> list of F ints in {-1, 0, 1}, or None."""
    if spec is None or spec == "" or spec == "()":
        return None
    if isinstance(spec, str):
        spec = [int(v) for v in spec.strip("()[] ").replace(" ", "").split(",") if v != ""]
    if isinstance(spec, dict):
        out = [0] * F
        for k, v in spec.items():
            i = feature_names.index(k) if (feature_names and k in feature_names) else int(str(k).lstrip("f"))
            out[i] = int(v)
        spec = out
    spec = [int(v) for v in spec]
    if len(spec) > F or any(v not in (-1, 0, 1) for v in spec):
        raise ValueError("monotone_constraints must hold at most %d values from {-1, 0, 1}" % F)
    spec = spec + [0] * (F - len(spec))
    return spec if any(spec) else None


def parse_interactions(spec, F):
    """``"[[0, 1], [2, 3, 4]]"`` / list of lists of feature indices -> list of frozensets, or None."""
    if spec is None or spec == "" or spec == "[]":
        return None
    if isinstance(spec, str):
        spec = json.loads(spec)
    groups = []
    for grp in spec:
        g = frozenset(int(str(v).lstrip("f")) for v in grp)
        if any(v < 0 or v >= F for v in g):
            raise ValueError("interaction_constraints refer to a feature outside [0, %d)" % F)
        groups.append(g)
    return groups or None


def _allowed_by_interactions(path, groups, F):
    """Features a node may split on, given the features already used on its path (xgboost's rule: the path
    features themselves plus every group that contains the whole path; everything at the root)."""
    if not path:
        return None
    ok = set(path)
    for g in groups:
        if path <= g:
            ok |= g
    return sorted(ok)


def _threshold_l1(g, alpha):
    if alpha <= 0:
        return g
    return torch.where(g > alpha, g - alpha, torch.where(g < -alpha, g + alpha, torch.zeros_like(g)))


def train(params, X, y, weight=None, eval_set=None, base_margin=None, eval_base_margin=None, missing=float("nan"),
          xgb_model=None, callbacks=None, device=None):
    """Train a :class:`Booster`.

    ``X``/``y``: this worker's shard (numpy or torch).  ``eval_set``: ``(X_val, y_val, w_val)`` or None.
    Inside a multi-worker gang (``hvd.init()`` done, size > 1) sketches, histograms and metrics are
    combined across workers, so every worker returns the identical model.
    Returns ``(booster, evals_result)``.
    """
    p = normalize_params(params)
    dist = _Dist()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if (torch.cuda.is_available() and params.get("use_gpu")) \
            else torch.device("cpu")
    X = torch.as_tensor(np.ascontiguousarray(X) if isinstance(X, np.ndarray) else X, dtype=torch.float32).to(device)
    y = torch.as_tensor(np.asarray(y) if not torch.is_tensor(y) else y, dtype=torch.float32).to(device).reshape(-1)
    n, F = X.shape
    w = torch.ones(n, device=device) if weight is None else torch.as_tensor(np.asarray(weight), dtype=torch.float32).to(device)
    objective = p["objective"] or "reg:squarederror"
    spw = float(p.get("scale_pos_weight") or 1.0)
    if spw != 1.0 and objective in ("binary:logistic", "reg:logistic"):
        w = torch.where(y > 0.5, w * spw, w)          # xgboost semantics: positives weigh scale_pos_weight times more
    non_default = [k for k in IGNORED_PARAMS if params.get(k) not in (None, DEFAULTS.get(k), 1, 1.0, "gbtree", "hist", "gain", 0.0)]
    if non_default and dist.rank == 0:
        import warnings
        warnings.warn("sparkdl.xgboost: parameters accepted for API compatibility but not used by the B200 GBDT engine: %s"
                      % ", ".join(non_default))
    num_class = int(p["num_class"] or 1)
    obj = _Objective(objective, num_class, float(p.get("max_delta_step") or 0.0), float(p.get("tweedie_variance_power") or 1.5))
    G = obj.n_groups
    rng = torch.Generator(device="cpu").manual_seed(int(p["random_state"] or 0) * 1000003 + dist.rank)  # rows: per shard
    rng_cols = torch.Generator(device="cpu").manual_seed(int(p["random_state"] or 0))                  # columns: same everywhere

    if xgb_model is not None:
        booster = xgb_model.copy() if isinstance(xgb_model, Booster) else Booster.load_model(xgb_model)
        booster.missing = missing
    else:
        booster = Booster(objective, num_class, float(p["base_score"]), missing, F)
    cuts = compute_cuts(X, missing, int(p["max_bin"]), dist, seed=int(p["random_state"] or 0))
    bins = bin_features(X, cuts, missing)
    builder = _HistBuilder(bins, dist, use_fused=os.environ.get("SPARKDL_GBDT_FUSED", "1") != "0")

    margin = booster.predict_margin(X, base_margin)
    ev = None
    if eval_set is not None and eval_set[0] is not None:
        Xv = torch.as_tensor(np.ascontiguousarray(eval_set[0]), dtype=torch.float32).to(device).reshape(-1, F)
        yv = torch.as_tensor(np.asarray(eval_set[1]), dtype=torch.float32).to(device).reshape(-1)
        wv = torch.ones(yv.shape[0], device=device) if eval_set[2] is None else \
            torch.as_tensor(np.asarray(eval_set[2]), dtype=torch.float32).to(device)
        ev = {"X": Xv, "y": yv, "w": wv, "margin": booster.predict_margin(Xv, eval_base_margin)}
    metrics = p["eval_metric"] or obj.default_metric()
    metrics = [metrics] if isinstance(metrics, str) else list(metrics)
    evals_result = {"validation_0": {m: [] for m in metrics}} if ev is not None else {}
    esr = p["early_stopping_rounds"]
    best, best_iter = None, -1
    lam, alpha, gamma, mcw = float(p["reg_lambda"]), float(p["reg_alpha"]), float(p["gamma"]), float(p["min_child_weight"])
    eta, max_depth = float(p["learning_rate"]), int(p["max_depth"])
    start_round = booster.num_boosted_rounds
    arange_n = torch.arange(n, device=device)
    # bin b of feature f is a split candidate iff b < number of cuts of f
    bin_has_cut = torch.arange(MISSING_BIN, device=device).view(1, 1, -1) < \
        torch.tensor([c.numel() for c in cuts], device=device).view(1, F, 1)
    mds = float(p.get("max_delta_step") or 0.0)
    if objective == "count:poisson" and mds == 0.0:
        mds = 0.7                                     # xgboost's safeguard for the exp link
    mono = parse_monotone(p.get("monotone_constraints"), F)
    mono_t = torch.tensor(mono, device=device, dtype=torch.float64).view(1, F, 1) if mono else None
    weighted_gain = mds > 0.0 or mono is not None    # gain from (clamped) child weights instead of the closed form
    inter = parse_interactions(p.get("interaction_constraints"), F)
    cs_level, cs_node = float(p.get("colsample_bylevel") or 1.0), float(p.get("colsample_bynode") or 1.0)
    INF = float("inf")

    def leaf_weight(Gx, Hx, lo=None, hi=None):
        wx = -_threshold_l1(Gx, alpha) / (Hx + lam)
        if mds > 0.0:
            wx = wx.clamp(-mds, mds)
        if lo is not None:
            wx = torch.maximum(torch.minimum(wx, hi), lo)
        return wx

    def gain_of_weight(Gx, Hx, wx):                   # -2 * (G w + (H + lambda) w^2 / 2 + alpha |w|)
        return -(2.0 * Gx * wx + (Hx + lam) * wx * wx + 2.0 * alpha * wx.abs())

    npt = max(1, int(p.get("num_parallel_tree") or 1))
    if xgb_model is not None and int(booster.attributes.get("num_parallel_tree", 1)) != npt and booster.trees:
        raise ValueError("xgb_model was trained with num_parallel_tree=%s" % booster.attributes.get("num_parallel_tree", 1))
    booster.attributes["num_parallel_tree"] = npt
    sibling_subtract = os.environ.get("SPARKDL_GBDT_SIBLING_SUBTRACT", "1") != "0"
    skip_leaf_hist = os.environ.get("SPARKDL_GBDT_SKIP_LEAF_HIST", "1") != "0"

    for it in range(int(p["n_estimators"])):
        g_all, h_all = obj.grad_hess(margin, y, w)
        # row subsampling (same decision stream on every worker given its own shard)
        if float(p["subsample"]) < 1.0:
            keep = (torch.rand(n, generator=rng) < float(p["subsample"])).to(device)
        else:
            keep = None
        new_trees, new_groups = [], []
        for k, j in [(k_, j_) for k_ in range(G) for j_ in range(npt)]:
            if j > 0 and float(p["subsample"]) < 1.0:        # every tree of a boosted forest draws its own rows
                keep = (torch.rand(n, generator=rng) < float(p["subsample"])).to(device)
            g, h = g_all[:, k], h_all[:, k]
            # fixed-point quantisation: |q| <= 2^30 per row keeps int64 sums exact for < 2^32 rows
            gmax, hmax = dist.allreduce_host([float(g.abs().max()) if n else 0.0, float(h.abs().max()) if n else 0.0], "max")
            sg = (2.0 ** 30) / max(gmax, 1e-30)
            sh = (2.0 ** 30) / max(hmax, 1e-30)
            gh_q = torch.stack([(g * sg).round(), (h * sh).round()], 1).to(torch.int32).contiguous()
            if float(p["colsample_bytree"]) < 1.0:
                kf = max(1, int(round(F * float(p["colsample_bytree"]))))
                fmask = torch.zeros(F, dtype=torch.bool)
                fmask[torch.randperm(F, generator=rng_cols)[:kf]] = True
                fmask = fmask.to(device)
            else:
                fmask = None
            tree = Tree()
            tree.add_node()
            pos = torch.zeros(n, dtype=torch.int32, device=device)       # slot within the current level
            if keep is not None:
                pos = torch.where(keep, pos, torch.full_like(pos, -1))
            node_of_row = torch.zeros(n, dtype=torch.long, device=device)  # tree node id (for leaf values)
            level_nodes = [0]
            node_lo, node_hi = {0: -INF}, {0: INF}      # monotone-constraint weight bounds per tree node
            node_path = {0: frozenset()}                # features used on the way to each node (interaction constraints)
            prev_int = None            # int64 histograms of the previous level (parents of this level's pairs)
            pair_parent, pair_build_left = [], []
            child_tot = None           # int64 [nodes, 2] totals of this level, known from the parents' splits
            for depth in range(max_depth + 1):
                n_nodes = len(level_nodes)
                if depth == max_depth and child_tot is not None and skip_leaf_hist:
                    hist_int, tot_int = None, child_tot                # leaf level: only the totals are needed
                elif prev_int is not None and sibling_subtract:
                    # Children come in (left, right) pairs.  Build only the lighter child (by hessian mass - known
                    # exactly and identically on every worker from the parent's histogram) and derive its sibling
                    # as parent - built: half the rows, half the atomics, half the allreduce bytes; exact in int64.
                    n_pairs = n_nodes // 2
                    build_slots = torch.tensor([2 * j + (0 if pair_build_left[j] else 1) for j in range(n_pairs)], device=device)
                    other_slots = torch.tensor([2 * j + (1 if pair_build_left[j] else 0) for j in range(n_pairs)], device=device)
                    lut = torch.full((n_nodes,), -1, dtype=torch.int32, device=device)
                    lut[build_slots] = torch.arange(n_pairs, dtype=torch.int32, device=device)
                    pos_build = torch.where(pos >= 0, lut[torch.clamp(pos, min=0).long()], torch.full_like(pos, -1))
                    built = builder.build(gh_q, pos_build, n_pairs)
                    hist_int = torch.empty((n_nodes, F, MAX_BINS, 2), dtype=torch.int64, device=device)
                    hist_int[build_slots] = built
                    hist_int[other_slots] = prev_int[torch.tensor(pair_parent, device=device)] - built
                    tot_int = hist_int[:, 0].sum(1)
                else:
                    hist_int = builder.build(gh_q, pos, n_nodes)
                    tot_int = hist_int[:, 0].sum(1)                    # [nodes, 2]: feature 0 sees every row once
                Gt, Ht = tot_int[:, 0].to(torch.float64) / sg, tot_int[:, 1].to(torch.float64) / sh
                lo_t = torch.tensor([node_lo[nid] for nid in level_nodes], device=device, dtype=torch.float64)
                hi_t = torch.tensor([node_hi[nid] for nid in level_nodes], device=device, dtype=torch.float64)
                w_node = leaf_weight(Gt, Ht, lo_t, hi_t) if weighted_gain else leaf_weight(Gt, Ht)
                vals_l, cover_l = (w_node * eta / npt).tolist(), Ht.tolist()      # one device->host copy per level
                for s, nid in enumerate(level_nodes):
                    tree.value[nid] = vals_l[s]
                    tree.cover[nid] = cover_l[s]
                if depth == max_depth:
                    break
                hist = hist_int.to(torch.float64)
                hist[..., 0] /= sg
                hist[..., 1] /= sh
                # split search: prefix sums over real bins, missing bin tried on both sides
                real = hist[:, :, :MISSING_BIN, :]
                miss = hist[:, :, MISSING_BIN, :]
                cs = torch.cumsum(real, 2)                               # left stats for "bin <= b"
                GL, HL = cs[..., 0], cs[..., 1]

                def score(Gx, Hx):
                    return _threshold_l1(Gx, alpha) ** 2 / (Hx + lam)
                parent = (gain_of_weight(Gt, Ht, w_node) if weighted_gain else score(Gt, Ht)).view(-1, 1, 1)
                # column sampling below the tree level: same random stream on every worker
                allowed = fmask if fmask is not None else torch.ones(F, dtype=torch.bool, device=device)
                if cs_level < 1.0:
                    cand = torch.nonzero(allowed.cpu())[:, 0]
                    kl = max(1, int(round(cand.numel() * cs_level)))
                    lvl = torch.zeros(F, dtype=torch.bool)
                    lvl[cand[torch.randperm(cand.numel(), generator=rng_cols)[:kl]]] = True
                    allowed = lvl.to(device)
                node_mask = None
                if cs_node < 1.0:
                    cand = torch.nonzero(allowed.cpu())[:, 0]
                    kn = max(1, int(round(cand.numel() * cs_node)))
                    nm = torch.zeros((n_nodes, F), dtype=torch.bool)
                    for s_ in range(n_nodes):
                        nm[s_, cand[torch.randperm(cand.numel(), generator=rng_cols)[:kn]]] = True
                    node_mask = nm.to(device).view(n_nodes, F, 1)
                if inter is not None and any(node_path[nid] for nid in level_nodes):
                    im = torch.ones((n_nodes, F), dtype=torch.bool)
                    for s_, nid in enumerate(level_nodes):
                        feats = _allowed_by_interactions(node_path[nid], inter, F)
                        if feats is not None:
                            im[s_] = False
                            im[s_, feats] = True
                    im = im.to(device).view(n_nodes, F, 1)
                    node_mask = im if node_mask is None else (node_mask & im)
                best_gain = None
                for dl in (1, 0):                                        # missing goes left / right
                    GLd = GL + (miss[..., 0:1] if dl else 0)
                    HLd = HL + (miss[..., 1:2] if dl else 0)
                    GRd = Gt.view(-1, 1, 1) - GLd
                    HRd = Ht.view(-1, 1, 1) - HLd
                    ok = (HLd >= mcw) & (HRd >= mcw)
                    if weighted_gain:
                        wl = leaf_weight(GLd, HLd, lo_t.view(-1, 1, 1), hi_t.view(-1, 1, 1))
                        wr = leaf_weight(GRd, HRd, lo_t.view(-1, 1, 1), hi_t.view(-1, 1, 1))
                        gain = 0.5 * (gain_of_weight(GLd, HLd, wl) + gain_of_weight(GRd, HRd, wr) - parent) - gamma
                        if mono_t is not None:
                            ok &= ~(((mono_t > 0) & (wl > wr)) | ((mono_t < 0) & (wl < wr)))
                    else:
                        gain = 0.5 * (score(GLd, HLd) + score(GRd, HRd) - parent) - gamma
                    ok &= bin_has_cut                                    # a cut must exist at this bin
                    if fmask is not None or cs_level < 1.0:
                        ok &= allowed.view(1, F, 1)
                    if node_mask is not None:
                        ok &= node_mask
                    gain = torch.where(ok, gain, torch.full_like(gain, -float("inf")))
                    flat = gain.view(n_nodes, -1)
                    gmax_, arg = flat.max(1)
                    if best_gain is None:
                        best_gain, best_arg, best_dl = gmax_, arg, torch.full_like(arg, dl)
                    else:
                        upd = gmax_ > best_gain
                        best_gain = torch.where(upd, gmax_, best_gain)
                        best_arg = torch.where(upd, arg, best_arg)
                        best_dl = torch.where(upd, torch.full_like(arg, dl), best_dl)
                bf_t, bb_t = best_arg // MISSING_BIN, best_arg % MISSING_BIN
                bf, bb = bf_t.tolist(), bb_t.tolist()
                bg = best_gain.tolist()
                bdl = best_dl.tolist()
                # exact (integer) totals of the two children of every node under its best split
                ar = torch.arange(n_nodes, device=device)
                sel = hist_int[ar, bf_t]                                 # [nodes, 256, 2]
                left_int = torch.cumsum(sel[:, :MISSING_BIN], 1)[ar, bb_t] + \
                    torch.where(best_dl.bool().view(-1, 1), sel[:, MISSING_BIN], torch.zeros_like(sel[:, MISSING_BIN]))
                right_int = tot_int - left_int
                left_lighter = (left_int[:, 1] <= right_int[:, 1]).tolist()
                if mono is not None:
                    wl_b = leaf_weight(left_int[:, 0].to(torch.float64) / sg, left_int[:, 1].to(torch.float64) / sh, lo_t, hi_t).tolist()
                    wr_b = leaf_weight(right_int[:, 0].to(torch.float64) / sg, right_int[:, 1].to(torch.float64) / sh, lo_t, hi_t).tolist()
                next_tot, pair_parent, pair_build_left = [], [], []
                next_nodes, slot_feat, slot_bin, slot_dl, slot_left, slot_right = [], [], [], [], [], []
                for s, nid in enumerate(level_nodes):
                    if bg[s] > 1e-12 and math.isfinite(bg[s]):
                        pair_parent.append(s)
                        pair_build_left.append(bool(left_lighter[s]))
                        next_tot.append(left_int[s])
                        next_tot.append(right_int[s])
                        l, r = tree.add_node(), tree.add_node()
                        tree.left[nid], tree.right[nid] = l, r
                        node_path[l] = node_path[r] = node_path[nid] | {bf[s]}
                        node_lo[l] = node_lo[r] = node_lo[nid]
                        node_hi[l] = node_hi[r] = node_hi[nid]
                        if mono is not None and mono[bf[s]] != 0:
                            mid = 0.5 * (wl_b[s] + wr_b[s])
                            if mono[bf[s]] > 0:
                                node_hi[l], node_lo[r] = min(node_hi[l], mid), max(node_lo[r], mid)
                            else:
                                node_lo[l], node_hi[r] = max(node_lo[l], mid), min(node_hi[r], mid)
                        tree.feature[nid] = bf[s]
                        tree.threshold[nid] = float(cuts[bf[s]][bb[s]])
                        tree.default_left[nid] = int(bdl[s])
                        tree.gain[nid] = float(bg[s])
                        slot_feat.append(bf[s]); slot_bin.append(bb[s]); slot_dl.append(bdl[s])
                        slot_left.append(len(next_nodes)); next_nodes.append(l)
                        slot_right.append(len(next_nodes)); next_nodes.append(r)
                    else:
                        slot_feat.append(0); slot_bin.append(0); slot_dl.append(1)
                        slot_left.append(-1); slot_right.append(-1)
                if not next_nodes:
                    break
                # partition rows of this level into the next level's slots
                sf = torch.tensor(slot_feat, device=device)
                sb = torch.tensor(slot_bin, device=device)
                sd = torch.tensor(slot_dl, device=device, dtype=torch.bool)
                sl = torch.tensor(slot_left, device=device, dtype=torch.int32)
                sr = torch.tensor(slot_right, device=device, dtype=torch.int32)
                nl = torch.tensor([tree.left[nid] for nid in level_nodes], device=device)
                nr = torch.tensor([tree.right[nid] for nid in level_nodes], device=device)
                act = pos >= 0
                ps = torch.clamp(pos, min=0).long()
                rb = bins[arange_n, sf[ps]].long()
                go_left = torch.where(rb == MISSING_BIN, sd[ps], rb <= sb[ps])
                new_pos = torch.where(go_left, sl[ps], sr[ps])
                split_here = act & (sl[ps] >= 0)
                node_of_row = torch.where(split_here, torch.where(go_left, nl[ps], nr[ps]), node_of_row)
                pos = torch.where(split_here, new_pos, torch.full_like(pos, -1))
                level_nodes = next_nodes
                prev_int = hist_int
                child_tot = torch.stack(next_tot, 0)
            # leaf assignment for the margin update.  Sub-sampled-out rows were never routed:
            # run them through the finished tree.
            if keep is not None:
                tmp = Booster("reg:squarederror", 1, 0.0, missing, F)
                tmp.trees, tmp.tree_group = [tree], [0]
                delta = tmp.predict_margin(X, zero_base=True)[:, 0]
            else:
                vals = torch.tensor(tree.value, device=device, dtype=torch.float32)
                delta = vals[node_of_row]
            margin[:, k] += delta
            new_trees.append(tree)
            new_groups.append(k)
            booster.trees.append(tree)
            booster.tree_group.append(k)
        booster._packed = None
        rnd = start_round + it
        if ev is not None:
            tmp = Booster(objective, num_class, 0.0, missing, F)
            tmp.trees, tmp.tree_group = new_trees, new_groups
            ev["margin"] += tmp.predict_margin(ev["X"], zero_base=True)
            scores = {}
            for m in metrics:
                if m == "auc":
                    # rank statistic: needs every worker's validation rows, and every worker must see the SAME value
                    # (early stopping is decided from it)
                    cols = torch.stack([ev["margin"][:, 0], ev["y"], ev["w"]], 1).contiguous()
                    if dist.size > 1:
                        cols = dist.allgather(cols)
                    scores[m] = _auc(cols[:, 0], cols[:, 1], cols[:, 2])
                else:
                    s_, n_ = _metric_sums(m, obj, ev["margin"], ev["y"], ev["w"])
                    s_, n_ = dist.allreduce_host([s_, n_], "sum")
                    scores[m] = _finish_metric(m, s_, n_)
                evals_result["validation_0"][m].append(scores[m])
            last = scores[metrics[-1]]
            maximize = metrics[-1] in ("auc",)
            improved = best is None or (last > best if maximize else last < best)
            if improved:
                best, best_iter = last, rnd
            if p["verbose"] and dist.rank == 0:
                print("[%d]\t%s" % (rnd, "\t".join("validation_0-%s:%.6f" % kv for kv in scores.items())), flush=True)
            if callbacks:
                stop = False
                for cb in callbacks:
                    if callable(cb) and cb({"iteration": rnd, "evals_result": evals_result, "booster": booster}):
                        stop = True
                if dist.size > 1:                      # a callback that fires on one worker stops all of them
                    stop = dist.allreduce_host([1.0 if stop else 0.0], "max")[0] > 0
                if stop:
                    break
            if esr and rnd - best_iter >= int(esr):
                break
        elif callbacks:
            stop = any([bool(callable(cb) and cb({"iteration": rnd, "evals_result": evals_result, "booster": booster}))
                        for cb in callbacks])
            if dist.size > 1:
                stop = dist.allreduce_host([1.0 if stop else 0.0], "max")[0] > 0
            if stop:
                break
    if ev is not None and best is not None:
        booster.best_iteration, booster.best_score = best_iter, best
    booster.attributes["hist_launches"] = builder.launches
    booster.attributes["workers"] = dist.size
    return booster, evals_result
