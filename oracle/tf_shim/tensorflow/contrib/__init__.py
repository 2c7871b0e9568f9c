"""placeholder package of the numpy TensorFlow stand-in (see ../__init__.py)."""


def __getattr__(name):  # anything not defined here is an attribute sink (tf.contrib.slim, ...)
    if name.startswith("__"):
        raise AttributeError(name)
    import tensorflow
    return tensorflow._Any()
