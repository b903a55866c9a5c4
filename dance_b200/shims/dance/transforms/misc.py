from dance_b200.transforms.misc import Compose, SetConfig  # noqa: F401
