"""tensorflow.contrib.image.python.ops.image_ops: imported by preprocessing/anchor_manipulator.py, unused on this path."""
