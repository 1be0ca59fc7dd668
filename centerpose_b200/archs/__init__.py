"""Backbone descriptions: parameter trees (reference-compatible ``state_dict`` names) and
their lowering to the fused op program."""
