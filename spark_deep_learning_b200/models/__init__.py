"""Model zoo for the headline configs of BASELINE.json: ResNet-50 and BERT-large."""
from spark_deep_learning_b200.models.resnet import resnet50, resnet18ish  # noqa: F401


def bert_large(**kw):
    from spark_deep_learning_b200.models.bert import bert_large as _b
    return _b(**kw)
