"""T0 - API contract.  Re-expresses the four assertions of
/root/reference/tests/horovod/runner_base_test.py:26-59 for py3.12 and adds the documented
contracts the reference never tested (log_to_driver truncation, verbosity validation)."""
import inspect
from inspect import FullArgSpec, getfullargspec

import pytest

import sparkdl
from sparkdl import HorovodRunner


def test_func_signature():
    init_spec = getfullargspec(HorovodRunner.__init__)
    assert init_spec == FullArgSpec(
        args=['self'], varargs=None, varkw=None, defaults=None,
        kwonlyargs=['np', 'driver_log_verbosity'],
        kwonlydefaults={'driver_log_verbosity': 'log_callback_only'}, annotations={})
    run_spec = getfullargspec(HorovodRunner.run)
    assert run_spec == FullArgSpec(
        args=['self', 'main'], varargs=None, varkw='kwargs', defaults=None, kwonlyargs=[],
        kwonlydefaults=None, annotations={})


def test_init_keyword_only():
    with pytest.raises(TypeError):
        HorovodRunner(2)  # pylint: disable=too-many-function-args,missing-kwoa


def test_public_surface():
    assert sparkdl.__all__ == ['HorovodRunner']
    assert sparkdl.__version__ == '2.2.0-db1'
    import sparkdl.horovod as sh
    assert sh.__all__ == ['log_to_driver']
    assert inspect.signature(sh.log_to_driver).parameters.keys() == {'message'}
    hr = HorovodRunner(np=-2, driver_log_verbosity="all")
    assert hr.num_processor == -2


def test_verbosity_validation():
    with pytest.raises(ValueError):
        HorovodRunner(np=-1, driver_log_verbosity="loud")
    with pytest.raises(TypeError):
        HorovodRunner(np="2")


def test_log_to_driver_truncates(capsys):
    from sparkdl.horovod import log_to_driver
    log_to_driver("x" * 5000)      # no launcher: the driver is this process's stdout
    out = capsys.readouterr().out.strip()
    assert len(out) == 4000
    with pytest.raises(TypeError):
        log_to_driver(123)


def test_run_inprocess_optin(monkeypatch):
    """The open-source stub runs main() in-process (reference test_run); opt-in here."""
    monkeypatch.setenv("SPARKDL_INPROCESS_SINGLE", "1")
    data = []
    HorovodRunner(np=-1).run(lambda value: data.append(value), value=1)
    assert data == [1]


def test_np_positive_exceeding_slots_fails(monkeypatch):
    monkeypatch.setenv("SPARKDL_NUM_GPUS", "2")
    with pytest.raises(RuntimeError, match="task slots"):
        HorovodRunner(np=3).run(lambda: 0)
    with pytest.raises(RuntimeError, match="task slots"):
        HorovodRunner(np=-3).run(lambda: 0)
