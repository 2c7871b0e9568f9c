from . import anchor_manipulator  # noqa: F401
