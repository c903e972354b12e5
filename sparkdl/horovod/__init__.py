"""Worker-side helpers of the HorovodRunner logging contract.

``log_to_driver`` is a stub that raises in the reference
(``/root/reference/sparkdl/horovod/__init__.py:20-25``); here it is implemented: the message
travels over the job's control socket and the driver prints it on stdout, truncated to
4000 characters.
"""
import os

_MAX_CHARS = 4000


def log_to_driver(message):
    """
    Send a log message (string type) to driver side, and driver will print log to stdout.
    If message length is greater than 4000, it will be truncated.
    """
    if not isinstance(message, str):
        raise TypeError("log_to_driver expects a str message, got %s" % type(message).__name__)
    message = message[:_MAX_CHARS]
    from spark_deep_learning_b200.parallel import context
    context.send_log_to_driver(message)


__all__ = ['log_to_driver']
