"""ORACLE (test infrastructure): seeded stand-in Florence-2 (see ``standin/florence.py``)."""
from standin.florence import *  # noqa: F401,F403
from standin.florence import (GEN, IMAGE_TOKEN, PROMPT_IDS, VOCAB, florence_config, florence_standin, input_ids_for,  # noqa: F401
                              pixel_values_from_u8)
