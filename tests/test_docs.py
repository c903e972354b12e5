"""docs/ stays consistent with the package (sphinx itself is not installable offline; ref docs/index.rst is the
authoritative list of public modules)."""
import importlib
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_conf_is_valid_python_and_names_the_project():
    ns = {"__file__": os.path.join(ROOT, "docs", "conf.py")}
    src = open(ns["__file__"]).read()
    exec(compile(src, ns["__file__"], "exec"), ns)          # pylint: disable=exec-used
    assert "sphinx.ext.autodoc" in ns["extensions"]
    assert ns.get("autoclass_content") == "both"


def test_every_automodule_imports_and_documents_its_public_names():
    text = open(os.path.join(ROOT, "docs", "index.rst")).read()
    mods = re.findall(r"^\.\. automodule:: (\S+)", text, flags=re.M)
    assert {"sparkdl", "sparkdl.horovod", "sparkdl.horovod.tensorflow.keras", "sparkdl.xgboost"} <= set(mods)
    for name in mods:
        m = importlib.import_module(name)
        for sym in getattr(m, "__all__", []):
            obj = getattr(m, sym)
            if callable(obj):
                assert (obj.__doc__ or "").strip(), "%s.%s has no docstring" % (name, sym)


def test_static_assets_referenced_by_the_template_exist():
    layout = open(os.path.join(ROOT, "docs", "_templates", "layout.html")).read()
    for asset in re.findall(r"_static/([\w.\-]+)", layout) or re.findall(r"static/([\w.\-]+)", layout):
        assert os.path.exists(os.path.join(ROOT, "docs", "static", asset)), asset


def test_generated_api_reference_is_complete():
    """docs/API.md is generated from the docstrings; no public name may be undocumented."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "docs", "gen_api.py")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-800:]
    text = open(os.path.join(ROOT, "docs", "API.md")).read()
    assert "(undocumented)" not in text and "HorovodRunner" in text and "allreduce" in text


def test_every_environment_knob_is_documented():
    """docs/CONFIG.md lists every SPARKDL_* variable the code reads."""
    used = set()
    for base in ("spark_deep_learning_b200", "sparkdl"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cc", ".cu", ".h", ".cuh")):
                    used |= set(re.findall(r"SPARKDL_[A-Z0-9_]+", open(os.path.join(dirpath, f), errors="ignore").read()))
    doc = open(os.path.join(ROOT, "docs", "CONFIG.md")).read()
    missing = sorted(v for v in used if v not in doc and not v.endswith("_"))
    assert not missing, missing
