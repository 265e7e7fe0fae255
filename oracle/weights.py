"""Back-compat alias: the synthetic weight / waveform generators live in ``synth.py`` at the repo root (they are inputs, not oracle code)."""
from synth import *  # noqa: F401,F403
from synth import _conformer, _dense, _ff  # noqa: F401
