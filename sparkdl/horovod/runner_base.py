"""HorovodRunner - launches a gang of GPU-pinned worker processes on one B200 box.

API parity with ``/root/reference/sparkdl/horovod/runner_base.py``: the constructor is
keyword-only ``(*, np, driver_log_verbosity="log_callback_only")`` (ref ``:39``) and
``run(self, main, **kwargs)`` (ref ``:76``); both carry no annotations because the reference
test-suite compares ``inspect.getfullargspec`` byte for byte
(``/root/reference/tests/horovod/runner_base_test.py:26-37``).

Where the reference only *documents* behaviour (its ``run`` calls ``main`` in-process, ref
``:97-103``), this class implements it; see :mod:`spark_deep_learning_b200.parallel.launcher`.
"""
import logging
import os

_VERBOSITY = ("all", "log_callback_only")


class HorovodRunner(object):
    """
    HorovodRunner runs distributed deep learning training jobs with a Horovod-style API.

    ``main`` is pickled with cloudpickle, shipped to ``|np|`` worker processes (one per GPU,
    or per CPU core on a CPU-only host), and executed there.  Inside ``main`` use
    ``import sparkdl.hvd as hvd`` (or ``horovod.torch`` through the bundled shim):
    ``hvd.init()``, ``hvd.allreduce``, ``hvd.DistributedOptimizer`` ... are backed by
    hand-written sm_100a kernels over NVLink peer memory, not by NCCL.
    """

    # pylint: disable=invalid-name
    def __init__(self, *, np, driver_log_verbosity="log_callback_only"):
        """
        :param np: number of parallel processes to use for the job.

            - If <0, spawn ``-np`` subprocesses on this (driver) node.  Training stdout and
              stderr go to the caller's stdout/stderr and to the driver log file.
            - If >0, "cluster" mode: ``np`` tasks start all together.  On a single box the
              local task slots (GPUs, else CPU cores) are the cluster; the job fails if ``np``
              exceeds them.
            - If 0 (deprecated), use every task slot.
        :param driver_log_verbosity: "all" or "log_callback_only" (default).  The first worker
            collects logs from all workers.  With "all" every worker line is streamed to the
            driver; with "log_callback_only" only messages sent through
            :func:`sparkdl.horovod.log_to_driver` (or a ``LogCallback``) are shown.  ``np<0``
            always streams, as documented for the local mode.
        """
        if driver_log_verbosity not in _VERBOSITY:
            raise ValueError("driver_log_verbosity must be one of %s, got %r"
                             % (_VERBOSITY, driver_log_verbosity))
        if not isinstance(np, int) or isinstance(np, bool):
            raise TypeError("np must be an int, got %r" % (np,))
        self.num_processor = np
        self.driver_log_verbosity = driver_log_verbosity

    def run(self, main, **kwargs):
        """
        Runs a training job invoking main(**kwargs) on every worker.

        Both the main function and the keyword arguments are serialized using cloudpickle and
        distributed to the workers.

        :param main: a Python function that contains the training code.
            The expected signature is `def main(**kwargs)` or compatible forms.  It runs in
            other processes: change global state inside the function and mind pickling limits.
        :param kwargs: keyword arguments passed to the main function at invocation time.
        :return: return value of the main function on rank 0 (must be cloudpickle-serializable).
        """
        if not callable(main):
            raise TypeError("main must be callable")
        from spark_deep_learning_b200.parallel import launcher
        from spark_deep_learning_b200.utils.config import get_config
        np_ = self.num_processor
        logger = logging.getLogger("HorovodRunner")

        if os.environ.get("SPARKDL_EXTERNAL_GANG") == "1":
            # Already inside a gang started by an external launcher (torchrun, a Spark barrier
            # stage): every task calls run(); join that gang instead of spawning a second one.
            world = int(os.environ.get("WORLD_SIZE", "1"))
            if abs(np_) not in (0, world):
                raise RuntimeError("HorovodRunner(np=%d) inside an external gang of %d tasks"
                                   % (np_, world))
            return main(**kwargs)

        if np_ > 0:
            # cluster mode: with a live SparkSession the gang is one Spark barrier stage (runner_base.py:54-61)
            from spark_deep_learning_b200.parallel import spark_gang
            spark = spark_gang.active_session()
            if spark is not None:
                logger.info("HorovodRunner: launching a Spark barrier job with %d tasks.", np_)
                return spark_gang.launch_on_spark(spark, main, kwargs, np_)
        num_proc = launcher.resolve_num_proc(np_)
        if num_proc == 1 and get_config(refresh=True).inprocess_single:
            # opt-in: behave like the open-source stub (same process, closures keep working)
            logger.warning("SPARKDL_INPROCESS_SINGLE=1: running main() inside the driver process.")
            return main(**kwargs)
        stream_all = np_ < 0 or self.driver_log_verbosity == "all"
        logger.info("HorovodRunner: launching %d worker process(es).", num_proc)
        return launcher.launch_gang(main, kwargs, num_proc, stream_all)
