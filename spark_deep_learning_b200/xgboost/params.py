"""Spark-ML style ``Param`` / ``Params`` / ``Estimator`` / ``Model`` machinery.

The reference estimators subclass ``pyspark.ml`` base classes
(``/root/reference/sparkdl/xgboost/xgboost.py:31-39,109,125``).  pyspark is not installable on
the B200 box (SURVEY.md §7 "Missing"), so when it is absent this module provides a compatible,
dependency-free implementation of the pieces the estimator contract needs: typed params with
converters, param maps, ``copy``/``extractParamMap`` (what ``Pipeline`` / ``CrossValidator`` rely on),
and ``MLReadable`` / ``MLWritable`` persistence.  When pyspark *is* importable its classes are used
instead, so the estimators plug into real Spark pipelines.
"""
import copy as _copy
import json
import os
import uuid

try:  # pragma: no cover - exercised only where pyspark exists
    from pyspark.ml import Estimator, Model  # noqa: F401
    from pyspark.ml.param import Param, Params, TypeConverters  # noqa: F401
    from pyspark.ml.param.shared import (HasFeaturesCol, HasLabelCol, HasPredictionCol,  # noqa: F401
                                         HasProbabilityCol, HasRawPredictionCol,
                                         HasValidationIndicatorCol, HasWeightCol)
    from pyspark.ml.util import MLReadable, MLWritable  # noqa: F401
    HAVE_PYSPARK = True
except ImportError:
    HAVE_PYSPARK = False

if not HAVE_PYSPARK:

    class TypeConverters(object):
        @staticmethod
        def identity(value):
            return value

        @staticmethod
        def toInt(value):
            if isinstance(value, bool) or not isinstance(value, (int, float)) and not hasattr(value, "__int__"):
                raise TypeError("Could not convert %r to int" % (value,))
            if float(value) != int(value):
                raise TypeError("Could not convert %r to int" % (value,))
            return int(value)

        @staticmethod
        def toFloat(value):
            if isinstance(value, bool) or not isinstance(value, (int, float)):
                raise TypeError("Could not convert %r to float" % (value,))
            return float(value)

        @staticmethod
        def toString(value):
            if not isinstance(value, str):
                raise TypeError("Could not convert %r to string" % (value,))
            return value

        @staticmethod
        def toBoolean(value):
            if not isinstance(value, bool):
                raise TypeError("Boolean Param requires value of type bool. Found %s." % type(value))
            return value

    class Param(object):
        """A param with self-contained documentation (same fields as ``pyspark.ml.param.Param``)."""

        def __init__(self, parent, name, doc, typeConverter=None):
            self.parent = parent.uid if isinstance(parent, Params) else parent
            self.name = str(name)
            self.doc = str(doc)
            self.typeConverter = TypeConverters.identity if typeConverter is None else typeConverter

        def _copy_new_parent(self, parent):
            p = _copy.copy(self)
            p.parent = parent.uid
            return p

        def __str__(self):
            return "%s__%s" % (self.parent, self.name)

        def __repr__(self):
            return "Param(parent=%r, name=%r, doc=%r)" % (self.parent, self.name, self.doc)

        def __hash__(self):
            return hash(str(self))

        def __eq__(self, other):
            return isinstance(other, Param) and self.parent == other.parent and self.name == other.name

    class Params(object):
        """Components that take parameters; params are class attributes copied per instance."""

        def __init__(self):
            self._paramMap = {}
            self._defaultParamMap = {}
            self._params = None
            self.uid = "%s_%s" % (type(self).__name__, uuid.uuid4().hex[:12])
            self._copy_class_params()

        @staticmethod
        def _dummy():
            d = Params.__new__(Params)
            d.uid = "undefined"
            return d

        def _copy_class_params(self):
            for name in dir(type(self)):
                attr = getattr(type(self), name, None)
                if isinstance(attr, Param):
                    setattr(self, name, attr._copy_new_parent(self))

        @property
        def params(self):
            if self._params is None:
                self._params = sorted(
                    [getattr(self, n) for n in dir(self)
                     if n != "params" and not isinstance(getattr(type(self), n, None), property)
                     and isinstance(getattr(self, n, None), Param)], key=lambda p: p.name)
            return self._params

        def hasParam(self, paramName):
            p = getattr(self, paramName, None)
            return isinstance(p, Param)

        def getParam(self, paramName):
            p = getattr(self, paramName, None)
            if not isinstance(p, Param):
                raise ValueError("Cannot find param with name %s." % paramName)
            return p

        def _resolveParam(self, param):
            if isinstance(param, Param):
                if param.parent != self.uid:
                    return self.getParam(param.name)
                return param
            return self.getParam(param)

        def isSet(self, param):
            return self._resolveParam(param) in self._paramMap

        def hasDefault(self, param):
            return self._resolveParam(param) in self._defaultParamMap

        def isDefined(self, param):
            return self.isSet(param) or self.hasDefault(param)

        def getOrDefault(self, param):
            p = self._resolveParam(param)
            if p in self._paramMap:
                return self._paramMap[p]
            if p in self._defaultParamMap:
                return self._defaultParamMap[p]
            raise KeyError("Param %s is not set and has no default" % p.name)

        def set(self, param, value):
            p = self._resolveParam(param)
            self._paramMap[p] = p.typeConverter(value) if value is not None else None
            return self

        def _set(self, **kwargs):
            for k, v in kwargs.items():
                p = self.getParam(k)
                if v is not None:
                    try:
                        v = p.typeConverter(v)
                    except TypeError as e:
                        raise TypeError('Invalid param value given for param "%s". %s' % (p.name, e))
                self._paramMap[p] = v
            return self

        def _setDefault(self, **kwargs):
            for k, v in kwargs.items():
                p = self.getParam(k)
                self._defaultParamMap[p] = p.typeConverter(v) if v is not None else None
            return self

        def clear(self, param):
            self._paramMap.pop(self._resolveParam(param), None)

        def extractParamMap(self, extra=None):
            m = dict(self._defaultParamMap)
            m.update(self._paramMap)
            if extra:
                m.update(extra)
            return m

        def explainParam(self, param):
            p = self._resolveParam(param)
            vals = []
            if self.hasDefault(p):
                vals.append("default: %s" % (self._defaultParamMap[p],))
            if self.isSet(p):
                vals.append("current: %s" % (self._paramMap[p],))
            return "%s: %s (%s)" % (p.name, p.doc, ", ".join(vals) if vals else "undefined")

        def explainParams(self):
            return "\n".join(self.explainParam(p) for p in self.params)

        def copy(self, extra=None):
            that = _copy.copy(self)
            that._paramMap = dict(self._paramMap)
            that._defaultParamMap = dict(self._defaultParamMap)
            that._params = None
            if extra:
                for p, v in extra.items():
                    that.set(p if isinstance(p, str) else p.name, v)
            return that

    def _shared(name, doc, default=None, conv=TypeConverters.toString):
        getter = "get" + name[0].upper() + name[1:]

        def init(self):
            Params.__init__(self)

        cls = type("Has" + name[0].upper() + name[1:], (Params,), {
            name: Param(Params._dummy(), name, doc, conv),
            getter: lambda self: self.getOrDefault(getattr(self, name)),
            "_shared_default": (name, default),
        })
        return cls

    HasFeaturesCol = _shared("featuresCol", "features column name.", "features")
    HasLabelCol = _shared("labelCol", "label column name.", "label")
    HasWeightCol = _shared("weightCol", "weight column name. If this is not set or empty, we treat all "
                           "instance weights as 1.0.")
    HasPredictionCol = _shared("predictionCol", "prediction column name.", "prediction")
    HasProbabilityCol = _shared("probabilityCol", "Column name for predicted class conditional probabilities.",
                                "probability")
    HasRawPredictionCol = _shared("rawPredictionCol", "raw prediction (a.k.a. confidence) column name.",
                                  "rawPrediction")
    HasValidationIndicatorCol = _shared(
        "validationIndicatorCol", "name of the column that indicates whether each row is for training or for "
        "validation. False indicates training; true indicates validation.")

    def _apply_shared_defaults(obj):
        for klass in type(obj).__mro__:
            sd = klass.__dict__.get("_shared_default")
            if sd and sd[1] is not None:
                obj._setDefault(**{sd[0]: sd[1]})

    class Estimator(Params):
        def fit(self, dataset, params=None):
            if params is None:
                return self._fit(dataset)
            if isinstance(params, (list, tuple)):
                return [self.copy(p)._fit(dataset) for p in params]
            if isinstance(params, dict):
                return self.copy(params)._fit(dataset)
            raise TypeError("Params must be either a param map or a list/tuple of param maps")

        def _fit(self, dataset):
            raise NotImplementedError()

    class Transformer(Params):
        def transform(self, dataset, params=None):
            if params is None:
                return self._transform(dataset)
            if isinstance(params, dict):
                return self.copy(params)._transform(dataset)
            raise TypeError("Params must be a param map")

        def _transform(self, dataset):
            raise NotImplementedError()

    class Model(Transformer):
        pass

    class _Writer(object):
        def __init__(self, instance):
            self.instance = instance
            self._overwrite = False

        def overwrite(self):
            self._overwrite = True
            return self

        def save(self, path):
            if os.path.exists(path):
                if not self._overwrite:
                    raise IOError("Path %s already exists. Use write().overwrite().save(path)." % path)
                import shutil
                shutil.rmtree(path)
            os.makedirs(path)
            self.instance._save_impl(path)

    class _Reader(object):
        def __init__(self, cls):
            self.cls = cls

        def load(self, path):
            return self.cls._load_impl(path)

    class MLWritable(object):
        def write(self):
            return _Writer(self)

        def save(self, path):
            self.write().save(path)

    class MLReadable(object):
        @classmethod
        def read(cls):
            return _Reader(cls)

        @classmethod
        def load(cls, path):
            return cls.read().load(path)
else:
    def _apply_shared_defaults(obj):  # pyspark's mixins set their own defaults
        return None

    from pyspark.ml import Transformer  # noqa: F401


def save_metadata(path, cls_name, uid, param_values, extra=None):
    meta = {"class": cls_name, "uid": uid, "paramMap": param_values, "extra": extra or {}}
    with open(os.path.join(path, "metadata.json"), "w") as f:
        json.dump(meta, f, indent=1, default=str)


def load_metadata(path):
    with open(os.path.join(path, "metadata.json")) as f:
        return json.load(f)
