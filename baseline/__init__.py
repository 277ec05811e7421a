"""Baseline arm: the unmodified reference run from ``baseline/_ref`` (see ``baseline/reference.py``)."""
