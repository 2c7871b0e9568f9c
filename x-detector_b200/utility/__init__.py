"""Evaluation helpers -- the surface of the reference's ``utility/`` package that the Light-Head R-CNN eval path uses."""
