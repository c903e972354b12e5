"""Sphinx configuration: autodoc of the four public modules, like the reference's docs/index.rst."""
import os
import sys

sys.path.insert(0, os.path.abspath(".."))
project = "sparkdl-b200"
author = "sparkdl-b200 authors"
release = os.environ.get("PACKAGE_VERSION", "0.1.0")
extensions = ["sphinx.ext.autodoc", "sphinx.ext.viewcode"]
autoclass_content = "both"
autodoc_mock_imports = ["tensorflow", "keras", "pyspark"]
html_theme = "alabaster"
exclude_patterns = ["_build"]
templates_path = ["_templates"]
html_static_path = ["static"]
