"""Test infrastructure: CPU oracle for the MERV fusion hot path (see fusion_oracle.py). Not part of the product."""
