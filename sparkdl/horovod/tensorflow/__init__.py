"""TensorFlow-side helpers (``/root/reference/sparkdl/horovod/tensorflow/__init__.py`` is empty too)."""
