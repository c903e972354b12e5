"""``np > 0``: the gang as ONE Spark barrier stage (``/root/reference/sparkdl/horovod/runner_base.py:54-61``).

"If ``np > 0``, this will launch a Spark job with ``np`` tasks starting all together and run the Horovod job on the
task nodes.  It will wait until ``np`` task slots are available to launch the job.  If ``np`` is greater than the total
number of task slots on the cluster, the job will fail."  That is Spark's barrier execution mode:
``sc.parallelize(range(np), np).barrier().mapPartitions(task).collect()`` - Spark's scheduler does the waiting and the
failing, ``BarrierTaskContext`` gives every task its rank (``partitionId``) and an ``allGather`` for the rendezvous.

Inside a task: rank 0 hosts the bootstrap store of this framework (the same unix-socket KV store the local launcher
uses; the collective runtime is single-host, so all tasks must land on one host - checked through ``getTaskInfos``) and
prints everything the workers send with ``log_to_driver`` to ITS stderr, i.e. "always merged into the first Spark
executor's stderr" (``runner_base.py:63-64``); the messages also travel back with the result and are printed on the
driver when the job ends (Spark offers no live channel from tasks to the driver).

pyspark is not installable on the B200 image; the code path is exercised against ``tests/fakes/pyspark`` (same API
surface: SparkSession / SparkContext.parallelize / RDD.barrier().mapPartitions / BarrierTaskContext).
"""
import os
import sys
import tempfile
import uuid


def active_session():
    """The active SparkSession, or None when pyspark is absent / no session exists."""
    try:
        from pyspark.sql import SparkSession
    except ImportError:
        return None
    try:
        return SparkSession.getActiveSession()
    except Exception:  # pylint: disable=broad-except
        return None


def total_slots(sc):
    """Task slots of the cluster (what the barrier scheduler compares ``np`` against)."""
    try:
        return int(sc._jsc.sc().maxNumConcurrentTasks())   # real Spark >= 3.1
    except Exception:  # pylint: disable=broad-except
        return int(sc.defaultParallelism)


def enter_gang(ctx, num_proc, tag="hvd"):
    """Called at the top of a barrier task: rendezvous + environment of one worker of the gang.

    Returns ``(rank, store_server_or_None, log_lines)``; the caller must call :func:`leave_gang` in a ``finally``."""
    from spark_deep_learning_b200.parallel.store import StoreServer
    rank = ctx.partitionId()
    hosts = [str(getattr(t, "address", t)).rsplit(":", 1)[0] for t in ctx.getTaskInfos()]
    if len(set(hosts)) > 1:
        raise RuntimeError("the sparkdl B200 runtime is single-host (NVLink / shared memory): the %d barrier tasks were "
                           "scheduled on %d hosts %s" % (num_proc, len(set(hosts)), sorted(set(hosts))))
    server, lines, msg = None, [], ""
    if rank == 0:
        job_dir = tempfile.mkdtemp(prefix="sparkdl-spark-")
        os.chmod(job_dir, 0o700)

        def sink(r, message):
            lines.append((r, message))
            sys.stderr.write("[%d] %s\n" % (r, message))     # first executor's stderr collects every worker's log
            sys.stderr.flush()
        server = StoreServer(os.path.join(job_dir, "store.sock"), log_sink=sink)
        msg = "%s|%s" % (job_dir, uuid.uuid4().hex[:12])
    job_dir, job_id = ctx.allGather(msg)[0].split("|")
    local_rank = sum(1 for h in hosts[:rank] if h == hosts[rank])
    os.environ.update({
        "SPARKDL_RANK": str(rank), "SPARKDL_SIZE": str(num_proc), "SPARKDL_LOCAL_RANK": str(local_rank),
        "SPARKDL_LOCAL_SIZE": str(sum(1 for h in hosts if h == hosts[rank])),
        "SPARKDL_STORE": os.path.join(job_dir, "store.sock"), "SPARKDL_JOB_ID": "%s-%s" % (tag, job_id),
        "SPARKDL_JOB_DIR": job_dir, "RANK": str(rank), "WORLD_SIZE": str(num_proc), "LOCAL_RANK": str(local_rank),
        "HOROVOD_RANK": str(rank), "HOROVOD_SIZE": str(num_proc)})
    os.environ.pop("SPARKDL_EXTERNAL_GANG", None)
    from spark_deep_learning_b200.parallel import context as _context
    _context.reset_context()          # a reused python worker must not keep the previous job's context
    return rank, server, lines


def leave_gang(ctx, server):
    """Orderly exit: tear the runtime down, then hold every task until all are done (rank 0 hosts the store)."""
    try:
        hvd = sys.modules.get("spark_deep_learning_b200.parallel.hvd")
        if hvd is not None and hvd.is_initialized():
            hvd.shutdown()
    finally:
        from spark_deep_learning_b200.parallel import context as _context
        _context.reset_context()
        ctx.barrier()
        if server is not None:
            job_dir = os.path.dirname(server.path) if getattr(server, "path", None) else None
            server.close()
            if job_dir and os.path.basename(job_dir).startswith("sparkdl-spark-"):
                import shutil
                shutil.rmtree(job_dir, ignore_errors=True)


def launch_on_spark(spark, main, kwargs, num_proc):
    """Run ``main(**kwargs)`` on ``num_proc`` barrier tasks; returns rank 0's return value."""
    import cloudpickle
    sc = spark.sparkContext
    slots = total_slots(sc)
    if num_proc > slots:
        raise RuntimeError("HorovodRunner(np=%d): only %d task slots are available on this cluster; the job fails as "
                           "documented." % (num_proc, slots))
    payload = cloudpickle.dumps((main, kwargs))

    def task(_rows):
        from pyspark import BarrierTaskContext
        import cloudpickle as cp
        ctx = BarrierTaskContext.get()
        rank, server, lines = enter_gang(ctx, num_proc)
        try:
            fn, kw = cp.loads(payload)
            result = fn(**kw)
        finally:
            leave_gang(ctx, server)
        if rank == 0:
            yield cp.dumps((result, lines))

    out = sc.parallelize(range(num_proc), num_proc).barrier().mapPartitions(task).collect()
    result, lines = cloudpickle.loads(out[0])
    for _r, message in lines:          # log_to_driver output reaches the driver's stdout (at job end)
        print(message, flush=True)
    return result
