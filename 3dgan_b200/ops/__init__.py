"""Layer API of the reference (`ops/` in algoterranean/3dgan), NHWC, on the B200 engine."""
