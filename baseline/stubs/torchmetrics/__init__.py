"""Minimal shim of the two torchmetrics functions the reference uses (utils/metrics.py:3); torchmetrics is not in this image."""
