"""placeholder package of the numpy TensorFlow stand-in (see ../__init__.py)."""
