"""Harness around the UNMODIFIED reference model (test / measurement infrastructure, not product code).

``baseline/_ref`` is a git-ignored copy of the reference's ``src/`` and ``configs/detrpose/`` made by
``baseline/vendor.py`` in the build container; it travels to the GPU box with the snapshot.  Nothing in
``detrpose_b200/`` imports from here.
"""
