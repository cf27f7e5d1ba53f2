"""Gen-2 layer API of the reference (`hem/ops/` in algoterranean/3dgan) on the B200 engine."""
from .layers import conv2d, deconv2d, dense, flatten, lrelu, rescale, rmse  # noqa: F401
