"""``horovod`` shim backed by the sm_100a runtime (not Uber's Horovod)."""
__sparkdl_shim__ = True
__version__ = "0.28.1+sparkdl.b200"
