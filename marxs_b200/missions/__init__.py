"""Instrument configurations built from marxs_b200 elements."""
