"""Drop-in ``dance`` namespace over dance_b200 (see dance_b200/dropin.py)."""
import logging

__b2_dropin__ = True
logger = logging.getLogger("dance")
if not logger.handlers:
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter("[%(levelname)s][%(asctime)s][%(name)s][%(funcName)s] %(message)s"))
    logger.addHandler(_h)
    logger.setLevel(logging.INFO)
