from typing import *  # noqa: F401,F403
from typing import Literal

NormMode = Literal["normalize", "standardize", "minmax", "l2"]
GeneSummaryMode = Literal["sum", "cv", "rv", "var"]
LogLevel = Literal["NOTSET", "DEBUG", "INFO", "WARNING", "ERROR"]
