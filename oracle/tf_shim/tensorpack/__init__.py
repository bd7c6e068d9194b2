"""Stand-in for the tensorpack names reference models.py:9-10 imports (training plumbing only)."""
