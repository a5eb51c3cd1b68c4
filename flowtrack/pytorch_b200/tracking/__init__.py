"""Tracking glue of the hot path (reference: lib/tracking/{net_utils,flow_utils}.py) + the joint per-frame-pair
pipeline that the bench and the multi-GPU driver use (pipeline.py)."""
