"""Empty stand-in so the reference's ``lib.metrics.utils`` import chain resolves
(lib/datasets/clustering.py:2).  No faiss functionality; test infrastructure only."""
