"""b200gan — B200-native training step for the algoterranean/3dgan conv models.

The directory is named `3dgan_b200`; import it as `b200gan` (the root-level shim `b200gan.py`
loads this directory under that module name, since a Python identifier cannot start with a digit).
"""
__version__ = "0.1.0"
