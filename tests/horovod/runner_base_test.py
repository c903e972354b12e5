"""Same file name and test names as ``/root/reference/tests/horovod/runner_base_test.py`` (``:26,39,44,55``), with the
py3.12-compatible assertions.  ``test_run`` documents the one deliberate deviation: ``np=-1`` spawns a real worker
process (the documented Databricks behaviour, ref ``runner_base.py:48``), so the captured-list side effect of the
open-source stub is only reproduced with ``SPARKDL_INPROCESS_SINGLE=1``."""
import os
import unittest
from inspect import FullArgSpec, getfullargspec

from sparkdl import HorovodRunner


class HorovodRunnerBaseTestCase(unittest.TestCase):

    def test_func_signature(self):
        init_spec = getfullargspec(HorovodRunner.__init__)
        self.assertEqual(init_spec, FullArgSpec(
            args=['self'], varargs=None, varkw=None, defaults=None,
            kwonlyargs=['np', 'driver_log_verbosity'],
            kwonlydefaults={'driver_log_verbosity': 'log_callback_only'}, annotations={}))
        run_spec = getfullargspec(HorovodRunner.run)
        self.assertEqual(run_spec, FullArgSpec(
            args=['self', 'main'], varargs=None, varkw='kwargs', defaults=None, kwonlyargs=[],
            kwonlydefaults=None, annotations={}))

    def test_init_keyword_only(self):
        with self.assertRaises(TypeError):
            HorovodRunner(2)  # pylint: disable=too-many-function-args,missing-kwoa

    def test_run(self):
        os.environ["SPARKDL_INPROCESS_SINGLE"] = "1"
        try:
            hr = HorovodRunner(np=-1)
            data = []

            def append(value):
                data.append(value)

            hr.run(append, value=1)
            self.assertEqual(data[0], 1)
        finally:
            del os.environ["SPARKDL_INPROCESS_SINGLE"]

    def test_return_value(self):
        hr = HorovodRunner(np=-1)
        return_value = hr.run(lambda: 42)
        self.assertEqual(return_value, 42)
