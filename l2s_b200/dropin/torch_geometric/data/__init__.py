from .batch import Batch, Data  # noqa: F401
