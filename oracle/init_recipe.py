"""Conditioned, seeded weight initialisation and synthetic images (moved to ``centerpose_b200/synth.py`` so that the
benchmark's product arm does not import anything from ``oracle/``); re-exported here for the oracle, the golden
generator and the tests."""
from centerpose_b200.synth import (HEAD_BIAS, HEAD_GAIN, OFFSET_GAIN, conditioned_state_dict, synth_images)  # noqa: F401
