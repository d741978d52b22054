"""Stub: toolbox.py:3 imports h5py at module top; the data loader is off the propagation path."""
