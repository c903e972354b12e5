import sparkdl  # noqa: F401  (import smoke test, like /root/reference/tests/__init__.py:15)
