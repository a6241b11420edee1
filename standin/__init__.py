"""Seeded stand-in checkpoints (NOT product, NOT a compute path of the product).

Real ``icon_detect_v3/model.pt`` and ``icon_caption_florence`` weights cannot be obtained offline, so every test, the
smoke run and both benchmark arms use seeded weights of the exact layer shapes.  This package only *defines and fills*
those checkpoints (plain ``torch.nn`` modules + a committed BatchNorm calibration):

* the B200 path consumes their ``state_dict`` (never their ``forward``);
* the CPU oracle (``oracle/``) runs their ``forward`` as the fp32 reference.

Nothing under ``omniparser_b200/`` imports this package.
"""
