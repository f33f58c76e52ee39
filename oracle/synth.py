"""ORACLE -- test infrastructure only.  Re-exports the synthetic-checkpoint generator (which lives in the
product package because bench.py needs it without touching oracle/)."""
from vidi_b200.synth import *  # noqa: F401,F403
from vidi_b200.synth import make_inputs, make_state_dict, make_tensor, tensor_specs  # noqa: F401
