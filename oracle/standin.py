"""ORACLE (test infrastructure): seeded stand-in detector (see ``standin/yolo_weights.py``)."""
from standin.yolo_weights import *  # noqa: F401,F403
from standin.yolo_weights import BN_CALIB, GOLDEN, yolo_standin  # noqa: F401
