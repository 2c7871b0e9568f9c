from . import resnet_v2, variables, xception_body  # noqa: F401
