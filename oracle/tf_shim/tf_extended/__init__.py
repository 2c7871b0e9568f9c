"""Placeholder for the `tf_extended` package that the reference's preprocessing/common_preprocessing.py imports but
does not ship (it comes from SSD-Tensorflow): attribute sink, nothing on the golden paths calls into it."""
from tensorflow import _Any


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Any()
