"""TEST INFRASTRUCTURE ONLY -- the seeded synthetic weights (reference state_dict schema)
used by the oracle and the golden generator.  The recipe lives in the product package
(`nope_b200/synth_weights.py`) because the bench and the tools need random-init weights too and
must not import oracle/."""
from nope_b200.synth_weights import *  # noqa: F401,F403
from nope_b200.synth_weights import (alias_encoder_keys, checksum, encoder_param_shapes,  # noqa: F401
                                     make_encoder_state_dict, make_full_state_dict,
                                     make_unet_state_dict, unet_param_shapes)
