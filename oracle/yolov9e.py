"""ORACLE (test infrastructure): the fp32 PyTorch YOLOv9-E restatement lives in ``standin/yolov9e.py`` (it doubles as the
definition of the seeded stand-in checkpoint); re-exported here for the oracle's callers."""
from standin.yolov9e import *  # noqa: F401,F403
from standin.yolov9e import YOLOv9E, export_torchscript  # noqa: F401
