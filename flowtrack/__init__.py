"""Namespace package: the product lives in flowtrack.pytorch_b200 (also reachable as the repo-root
directory name `flowtrack.pytorch_b200/`, a symlink)."""
