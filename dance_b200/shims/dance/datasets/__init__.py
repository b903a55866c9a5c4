from dance_b200.datasets import *  # noqa: F401,F403
