"""Shim for the older ``biahub.analysis`` namespace (``scripts/measure_psf.py:15-17``)."""
