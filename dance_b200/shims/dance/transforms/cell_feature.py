from dance_b200.transforms.cell_feature import *  # noqa: F401,F403
