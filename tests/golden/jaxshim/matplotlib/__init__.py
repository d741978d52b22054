"""Stub: the reference imports matplotlib at module top (wave_optics.py:5); plotting is never called by the fixtures."""
